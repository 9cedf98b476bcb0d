"""Host NumPy versions of the reference's small math utilities (utils/lie_algebra.py,
utils/derivatives.py, utils/transformations.py) and reprojection factor callables
(gbp/factors/reprojection.py), kept so that code written against those modules imports and runs.
The BA sweep itself evaluates the same formulas in CUDA (csrc/gbp_math.cuh)."""
import numpy as np

_EPS = np.finfo(float).eps


def S03_hat_operator(x):
    a, b, c = x[0], x[1], x[2]
    return np.array([[0., -c, b], [c, 0., -a], [-b, a, 0.]])


def so3exp(w):
    """utils/lie_algebra.py:32-42"""
    w = np.asarray(w, dtype=float)
    th = np.linalg.norm(w)
    if th < 3 * _EPS:
        return np.eye(3)
    W = S03_hat_operator(w)
    return np.eye(3) + (np.sin(th) / th) * W + ((1 - np.cos(th)) / th ** 2) * (W @ W)


def so3log(R):
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    return th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def proj(x):
    """utils/transformations.py:5-9"""
    x = np.asarray(x)
    if x.ndim == 1:
        return x[:-1] / x[-1]
    return (x[:, :-1].T / x[:, -1]).T


def proj_derivative(x):
    """utils/derivatives.py:48-50"""
    x = np.asarray(x, dtype=float)
    n = len(x) - 1
    out = np.zeros((n, n + 1))
    out[:, :n] = np.eye(n) / x[-1]
    out[:, n] = -x[:-1] / x[-1] ** 2
    return out


def dR_wx_dw(w, x):
    """utils/derivatives.py:36-45"""
    w = np.asarray(w, dtype=float)
    R = so3exp(w)
    inner = (np.outer(w, w) + (R.T - np.eye(3)) @ S03_hat_operator(w)) / (w @ w)
    return -(R @ S03_hat_operator(x)) @ inner


def jac_fd(inp, meas_fn, *args, delta=1e-8):
    """utils/derivatives.py:9-22  forward finite differences"""
    inp = np.asarray(inp, dtype=float)
    z0 = np.atleast_1d(meas_fn(inp, *args))
    jac = np.zeros((len(z0), len(inp)))
    for i in range(len(inp)):
        d = inp.copy()
        d[i] += delta
        jac[:, i] = (np.atleast_1d(meas_fn(d, *args)) - z0) / delta
    return jac


def check_jac(jac_fn, inp, meas_fn, *args, threshold=1e-3):
    """utils/derivatives.py:25-33"""
    diff = np.max(jac_fn(inp, *args) - jac_fd(inp, meas_fn, *args))
    if diff < threshold:
        print(f"Passed! Jacobian correct to within {threshold}")
    else:
        print(f"Failed: Jacobian difference to finite difference Jacobian not within threshold ({threshold})"
              f"\nMaximum discrepancy between Jacobian and finite diff Jacobian: {diff}")


def getT_axisangle(x):
    """utils/transformations.py:14-22"""
    T = np.eye(4)
    T[:3, :3] = so3exp(x[3:6])
    T[:3, 3] = x[0:3]
    return T


def reprojection_meas_fn(inp, K):
    """gbp/factors/reprojection.py:12-24"""
    inp = np.asarray(inp, dtype=float)
    assert len(inp) == 9
    return proj(K @ (so3exp(inp[3:6]) @ inp[6:9] + inp[:3]))


def reprojection_jac_fn(inp, K):
    """gbp/factors/reprojection.py:27-44"""
    inp = np.asarray(inp, dtype=float)
    assert len(inp) == 9
    R = so3exp(inp[3:6])
    JpK = proj_derivative(K @ (R @ inp[6:9] + inp[:3])) @ K
    return np.hstack([JpK, JpK @ dR_wx_dw(inp[3:6], inp[6:9]), JpK @ R])
