"""gbp_b200: B200-native Gaussian Belief Propagation for bundle adjustment.

Public surface
  gbp_b200.ba.create_ba_graph / BAFactorGraph   device-resident mirror of the reference's gbp_ba API
  gbp_b200.engine.BAEngine                      object wrapper of the C ABI (include/gbp_b200.h)
  gbp_b200.hostgraph                            generic host FactorGraph (reference config 1 plumbing)
  gbp_b200.compat/{gbp,utils,vis}               reference-named packages (put on sys.path by gbp_b200.run)
"""
__version__ = "0.1.0"
