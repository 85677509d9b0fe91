# Drop-in for simu_sff/image_warp.py (and its copies under sff_scripts_*/utils/) of ssTEM-restoration.
from sstem_restoration_b200.warp import image_warp  # noqa: F401
