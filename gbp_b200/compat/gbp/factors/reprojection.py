"""``gbp.factors.reprojection`` (reference: gbp/factors/reprojection.py:12-54)."""
import numpy as np
from gbp_b200.se3 import reprojection_meas_fn as meas_fn, reprojection_jac_fn as jac_fn, check_jac  # noqa: F401

if __name__ == "__main__":
    K = np.array([[517.306408, 0., 318.64304], [0., 516.469215, 255.313989], [0., 0., 1.]])
    check_jac(jac_fn, np.random.rand(9), meas_fn, K)
