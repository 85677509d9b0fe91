# Drop-in for the numerical functions of simu_sff/simuSFF.py of ssTEM-restoration (the PNG I/O wrapper
# `SimuSFF` stays with the caller: it needs skimage).  Same arguments, same `random` draws, bit-equal results.
from sstem_restoration_b200.sff_sim import cal_distance, get_two_points, degradation, noise, simu_sff  # noqa: F401
