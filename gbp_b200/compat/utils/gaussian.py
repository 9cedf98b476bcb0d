from gbp_b200.gaussian import NdimGaussian  # noqa: F401
