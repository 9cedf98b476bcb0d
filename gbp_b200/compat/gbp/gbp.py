"""``gbp.gbp``: FactorGraph / VariableNode / Factor with the reference's signatures (gbp/gbp.py)."""
from gbp_b200.hostgraph import FactorGraph, VariableNode, Factor  # noqa: F401
from gbp_b200.gaussian import NdimGaussian  # noqa: F401
