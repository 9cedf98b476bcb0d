"""``gbp.gbp_ba``: the bundle-adjustment graph on the B200 engine (reference: gbp/gbp_ba.py)."""
from gbp_b200.ba import (BAFactorGraph, FrameVariableNode, LandmarkVariableNode,  # noqa: F401
                         ReprojectionFactor, create_ba_graph)
