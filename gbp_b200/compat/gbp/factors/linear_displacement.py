"""``gbp.factors.linear_displacement``: h(x1, x2) = x2 - x1 (reference: gbp/factors/linear_displacement.py:8-14)."""
import numpy as np


def jac_fn(x):
    n = len(x) // 2
    return np.concatenate([-np.eye(n), np.eye(n)], axis=1)


def meas_fn(x):
    n = len(x) // 2
    x = np.asarray(x)
    return x[n:2 * n] - x[:n]
