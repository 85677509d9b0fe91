# Drop-in for simu_sff/flow_synthesis.py of ssTEM-restoration (gen_line, gen_flow): the distance field is
# evaluated by the sm_100a kernel, results are bit-equal numpy arrays.
from sstem_restoration_b200.sff_sim import gen_line, gen_flow  # noqa: F401
