# Drop-in for libs/sepconv/SeparableConvolution.py of ssTEM-restoration:
# same class, same .apply(input, vertical, horizontal), sm_100a kernels underneath.
from sstem_restoration_b200.sepconv import SeparableConvolution  # noqa: F401
