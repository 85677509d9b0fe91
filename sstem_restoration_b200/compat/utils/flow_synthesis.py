# Drop-in for sff_scripts_{unfolding,fusion}/utils/flow_synthesis.py of ssTEM-restoration: gen_flow returns
# (flow, flow2, mask) as the data providers expect (data_provider.py:223).
from sstem_restoration_b200.sff_sim import gen_line  # noqa: F401
from sstem_restoration_b200 import sff_sim as _sff


def gen_flow(height, width, k, b, line_width=5, fold_width=10, dis_k=0.1):
    return _sff.gen_flow(height, width, k, b, line_width, fold_width, dis_k, two_flows=True)
