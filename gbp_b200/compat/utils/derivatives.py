from gbp_b200.se3 import jac_fd, check_jac, dR_wx_dw, proj_derivative  # noqa: F401
