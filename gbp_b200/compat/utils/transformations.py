from gbp_b200.se3 import proj, getT_axisangle  # noqa: F401
