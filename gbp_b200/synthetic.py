"""Synthetic BAL problems (BASELINE configs 4-5; the reference has no generator).

Spec: intrinsics of fr1desk; C keyframes on a circle of radius 4 m
looking at the origin; L landmarks ~ U([-1,1]^3); every landmark observed by `obs_per_lmk`
distinct keyframes drawn uniformly among those that see it inside the 640x480 image with depth
> 0.5 m; z = projection + N(0, 2^2) px; initial keyframe translations perturbed by N(0, 0.02^2) m,
rotations exact, landmarks perturbed by N(0, 0.05^2) m; measurements emitted camera-major.
Everything is vectorised NumPy (10 M measurements in a few seconds).
"""
from __future__ import annotations

import numpy as np

from .balio import BALProblem

FR1_K4 = np.array([517.306408, 516.469215, 318.64304, 255.313989])


def _rodrigues(w):
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    W = np.zeros(w.shape[:-1] + (3, 3))
    W[..., 0, 1], W[..., 0, 2] = -w[..., 2], w[..., 1]
    W[..., 1, 0], W[..., 1, 2] = w[..., 2], -w[..., 0]
    W[..., 2, 0], W[..., 2, 1] = -w[..., 1], w[..., 0]
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * (W @ W)


def _log_so3(R):
    """Axis-angle of rotation matrices with angle in (0, pi)."""
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    return v * (th / (2.0 * np.sin(th)))[..., None]


def _look_at(ang, radius, height):
    n = len(ang)
    centre = np.stack([radius * np.cos(ang), radius * np.sin(ang), np.full(n, height)], axis=-1)
    fwd = -centre / np.linalg.norm(centre, axis=-1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=-1, keepdims=True)
    down = np.cross(fwd, right)
    R_cw = np.stack([right, down, fwd], axis=-2)          # rows = camera axes in world coords
    t_cw = -np.einsum("cij,cj->ci", R_cw, centre)
    return R_cw, t_cw


def look_at_cameras(n_cams, radius=4.0, height=0.5, max_angle=3.0):
    """T_cw = (R_cw, t_cw) of cameras on a circle looking at the origin (z forward, y down).

    Azimuths are spread evenly over the part of the circle where the axis-angle norm of R_cw stays
    in [0.05, max_angle] (the reference's dR_wx_dw divides by w.w and the log map degenerates at pi)."""
    grid = 2.0 * np.pi * (np.arange(36000) + 0.5) / 36000
    Rg, _ = _look_at(grid, radius, height)
    th = np.arccos(np.clip((np.trace(Rg, axis1=-2, axis2=-1) - 1.0) / 2.0, -1.0, 1.0))
    ok = grid[(th >= 0.05) & (th <= max_angle)]
    ang = ok[np.floor((np.arange(n_cams) + 0.5) * len(ok) / n_cams).astype(int)]
    return _look_at(ang, radius, height)


def make_synthetic(n_cams=1000, n_lmks=1_000_000, obs_per_lmk=10, seed=0, pixel_noise=2.0, cam_t_noise=0.02,
                   lmk_noise=0.05, K4=FR1_K4, chunk=200_000) -> BALProblem:
    rng = np.random.default_rng(seed)
    R_cw, t_cw = look_at_cameras(n_cams)
    w_cw = _log_so3(R_cw)
    nrm = np.linalg.norm(w_cw, axis=-1)
    assert nrm.min() >= 0.05 and nrm.max() <= 3.0 + 1e-9, (nrm.min(), nrm.max())
    R_cw = _rodrigues(w_cw)                                 # what the factor model will reconstruct
    lmks = rng.uniform(-1.0, 1.0, size=(n_lmks, 3))
    fx, fy, cx, cy = K4
    obs_per_lmk = min(obs_per_lmk, n_cams)
    cam_ids = np.empty((n_lmks, obs_per_lmk), dtype=np.int32)
    zs = np.empty((n_lmks, obs_per_lmk, 2))
    for s in range(0, n_lmks, chunk):
        y = lmks[s:s + chunk]
        n = len(y)
        # draw candidate cameras, keep the first obs_per_lmk distinct visible ones
        n_try = min(n_cams, max(4 * obs_per_lmk, 16))
        if n_try == n_cams:
            cand = np.argsort(rng.random((n, n_cams)), axis=1).astype(np.int32)
        else:
            cand = rng.integers(0, n_cams, size=(n, n_try)).astype(np.int32)
        pc = np.einsum("nkij,nj->nki", R_cw[cand], y) + t_cw[cand]
        u = fx * pc[..., 0] / pc[..., 2] + cx
        v = fy * pc[..., 1] / pc[..., 2] + cy
        ok = (pc[..., 2] > 0.5) & (u >= 0) & (u < 640) & (v >= 0) & (v < 480)
        # mask duplicates (same camera drawn twice for one landmark)
        order = np.argsort(cand, axis=1, kind="stable")
        sc = np.take_along_axis(cand, order, axis=1)
        dup_sorted = np.zeros_like(ok)
        dup_sorted[:, 1:] = sc[:, 1:] == sc[:, :-1]
        dup = np.zeros_like(ok)
        np.put_along_axis(dup, order, dup_sorted, axis=1)
        ok &= ~dup
        rank = np.cumsum(ok, axis=1)
        if (rank[:, -1] < obs_per_lmk).any():
            raise RuntimeError("synthetic generator: a landmark is visible from too few sampled cameras")
        sel = ok & (rank <= obs_per_lmk)
        idx = np.nonzero(sel)
        cam_ids[s:s + n] = cand[idx].reshape(n, obs_per_lmk)
        zs[s:s + n, :, 0] = u[idx].reshape(n, obs_per_lmk)
        zs[s:s + n, :, 1] = v[idx].reshape(n, obs_per_lmk)
    zs += rng.normal(0.0, pixel_noise, size=zs.shape)
    lmk_ids = np.repeat(np.arange(n_lmks, dtype=np.int32)[:, None], obs_per_lmk, axis=1)
    # camera-major emission like the reference data files
    flat_cam, flat_lmk, flat_z = cam_ids.ravel(), lmk_ids.ravel(), zs.reshape(-1, 2)
    order = np.argsort(flat_cam, kind="stable")
    cam0 = np.concatenate([t_cw + rng.normal(0.0, cam_t_noise, size=t_cw.shape), w_cw], axis=1)
    lmk0 = lmks + rng.normal(0.0, lmk_noise, size=lmks.shape)
    return BALProblem(flat_cam[order], flat_lmk[order], flat_z[order], cam0, lmk0, np.asarray(K4, dtype=np.float64))
