# Drop-in for simu_sff/image_warp.py of ssTEM-restoration.
from sstem_restoration_b200.warp import image_warp  # noqa: F401
