# Drop-in for sff_scripts_{unfolding,fusion}/utils/image_warp_torch.py of ssTEM-restoration.
from sstem_restoration_b200.warp import SpatialTransformation  # noqa: F401
