from gbp_b200.se3 import S03_hat_operator, so3exp, so3log  # noqa: F401
