"""CPU oracle for the GBP bundle-adjustment sweep  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement (float64, vectorised over factors instead of one
Python object per factor) of the reference algorithm in joeaortiz/gbp.  It exists to
CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``gbp_b200`` never does (it fails loudly when the CUDA library is missing).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function below
against fixtures produced by running the unmodified reference in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

It deliberately keeps the reference's arithmetic *form* - the explicit 9x9 factor
(eta_f, Lambda_f) and Schur complements through ``np.linalg.inv`` - whereas the CUDA
kernels use an algebraically equal low-rank (Woodbury) form.  Agreement between the two
is therefore a real check, not a tautology.

Every function cites the reference file:line it follows (paths under /root/reference).
"""
from __future__ import annotations

import numpy as np

_EPS = np.finfo(float).eps


# --------------------------------------------------------------------------- L0 math
def hat(x):
    """utils/lie_algebra.py:11-17  S03_hat_operator, batched: x[...,3] -> [...,3,3]."""
    x = np.asarray(x, dtype=np.float64)
    o = np.zeros(x.shape[:-1] + (3, 3))
    o[..., 0, 1] = -x[..., 2]
    o[..., 0, 2] = x[..., 1]
    o[..., 1, 0] = x[..., 2]
    o[..., 1, 2] = -x[..., 0]
    o[..., 2, 0] = -x[..., 1]
    o[..., 2, 1] = x[..., 0]
    return o


def so3exp(w):
    """utils/lie_algebra.py:32-42  Rodrigues formula; identity when theta < 3*eps."""
    w = np.asarray(w, dtype=np.float64)
    theta = np.linalg.norm(w, axis=-1)
    small = theta < _EPS * 3
    th = np.where(small, 1.0, theta)
    W = hat(w)
    a = (np.sin(th) / th)[..., None, None]
    b = ((1 - np.cos(th)) / th ** 2)[..., None, None]
    R = np.eye(3) + a * W + b * (W @ W)
    R[small] = np.eye(3)
    return R


def dR_wx_dw(w, x):
    """utils/derivatives.py:36-45  d(R(w)x)/dw = -R x^ (w w^T + (R^T - I) w^) / (w.w).

    Like the reference this is NaN at w == 0 (division by w.w)."""
    R = so3exp(w)
    ww = np.einsum("...i,...j->...ij", w, w)
    inner = (ww + (np.swapaxes(R, -1, -2) - np.eye(3)) @ hat(w)) / np.sum(w * w, axis=-1)[..., None, None]
    return -(R @ hat(x)) @ inner


def proj(p):
    """utils/transformations.py:5-7."""
    return p[..., :2] / p[..., 2:3]


def proj_derivative(p):
    """utils/derivatives.py:48-50  [[1/z,0,-x/z^2],[0,1/z,-y/z^2]]."""
    o = np.zeros(p.shape[:-1] + (2, 3))
    o[..., 0, 0] = 1.0 / p[..., 2]
    o[..., 1, 1] = 1.0 / p[..., 2]
    o[..., 0, 2] = -p[..., 0] / p[..., 2] ** 2
    o[..., 1, 2] = -p[..., 1] / p[..., 2] ** 2
    return o


def K_matrix(K4):
    fx, fy, cx, cy = [float(v) for v in K4]
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])


# --------------------------------------------------------------------------- L1 factor model
def meas_fn(x, K):
    """gbp/factors/reprojection.py:12-24   x[...,9] = [t(3), w(3), y(3)] -> pixel [...,2]."""
    R = so3exp(x[..., 3:6])
    pc = np.einsum("...ij,...j->...i", R, x[..., 6:9]) + x[..., 0:3]
    return proj(pc @ K.T)


def jac_fn(x, K):
    """gbp/factors/reprojection.py:27-44   2x9 Jacobian [Jp K | Jp K dRy/dw | Jp K R]."""
    w, y = x[..., 3:6], x[..., 6:9]
    R = so3exp(w)
    pc = np.einsum("...ij,...j->...i", R, y) + x[..., 0:3]
    JpK = proj_derivative(pc @ K.T) @ K
    J = np.zeros(x.shape[:-1] + (2, 9))
    J[..., 0:3] = JpK
    J[..., 3:6] = JpK @ dR_wx_dw(w, y)
    J[..., 6:9] = JpK @ R
    return J


# --------------------------------------------------------------------------- the graph
class BAOracle:
    """Structure-of-arrays restatement of BAFactorGraph (gbp/gbp_ba.py:12-69) and the
    FactorGraph sweep (gbp/gbp.py:46-92) for reprojection factors.

    Factor order = the reference's (gbp/gbp_ba.py:128-143): camera-major, file order
    within a camera, i.e. a stable sort of the measurement list by camera id.
    """

    def __init__(self, cam_id, lmk_id, z, cam0, lmk0, K4, configs):
        cam_id = np.asarray(cam_id, dtype=np.int64)
        order = np.argsort(cam_id, kind="stable")
        self.file_order = order
        self.cam = cam_id[order]
        self.lmk = np.asarray(lmk_id, dtype=np.int64)[order]
        self.z = np.asarray(z, dtype=np.float64)[order]
        self.C, self.L, self.F = len(cam0), len(lmk0), len(self.cam)
        self.K = K_matrix(K4)
        self.K4 = np.asarray(K4, dtype=np.float64)
        c = configs
        self.eta_damping = float(c["eta_damping"])           # gbp/gbp.py:28
        self.beta = float(c["beta"])                          # gbp/gbp.py:32
        self.num_undamped_iters = int(c["num_undamped_iters"])
        self.min_linear_iters = int(c["min_linear_iters"])
        self.var0 = float(c["gauss_noise_std"]) ** 2          # gbp/gbp.py:236
        self.loss = c.get("loss", None)
        self.Nstds = float(c.get("Nstds", 3.0))               # mahalanobis_threshold
        F = self.F
        # variable nodes (gbp/gbp.py:156-174, gbp/gbp_ba.py:114-125)
        self.cam_mu = np.array(cam0, dtype=np.float64)
        self.lmk_mu = np.array(lmk0, dtype=np.float64)
        self.cam_prior_eta = np.zeros((self.C, 6)); self.cam_prior_lam = np.zeros((self.C, 6, 6))
        self.lmk_prior_eta = np.zeros((self.L, 3)); self.lmk_prior_lam = np.zeros((self.L, 3, 3))
        self.cam_eta = np.zeros((self.C, 6)); self.cam_lam = np.zeros((self.C, 6, 6))
        self.lmk_eta = np.zeros((self.L, 3)); self.lmk_lam = np.zeros((self.L, 3, 3))
        # factor nodes (gbp/gbp.py:202-249): zero messages, damping 0, iters_since_relin 1
        self.msg_cam_eta = np.zeros((F, 6)); self.msg_cam_lam = np.zeros((F, 6, 6))
        self.msg_lmk_eta = np.zeros((F, 3)); self.msg_lmk_lam = np.zeros((F, 3, 3))
        self.adaptive_var = np.full(F, self.var0)
        self.robust_flag = np.zeros(F, dtype=bool)
        self.factor_damping = np.zeros(F)
        self.iters_since_relin = np.ones(F, dtype=np.int64)
        # gbp/gbp_ba.py:136-137: linearise every factor at the initial means
        self.linpoint = np.concatenate([self.cam_mu[self.cam], self.lmk_mu[self.lmk]], axis=1)
        self.factor_eta = np.zeros((F, 9)); self.factor_lam = np.zeros((F, 9, 9))
        self.compute_factor(np.ones(F, dtype=bool), self.linpoint)

    # ---- gbp/gbp.py:267-294
    def compute_factor(self, mask, linpoint):
        if not mask.any():
            return
        x0 = linpoint[mask]
        J = jac_fn(x0, self.K)
        h0 = meas_fn(x0, self.K)
        inv_var = 1.0 / self.adaptive_var[mask]
        JT = np.swapaxes(J, -1, -2)
        # lambda = J^T (I/var) J ; eta = (J^T (I/var)) (J x0 + z - h0)
        JTw = JT * inv_var[:, None, None]
        self.factor_lam[mask] = JTw @ J
        rhs = np.einsum("fij,fj->fi", J, x0) + self.z[mask] - h0
        self.factor_eta[mask] = np.einsum("fij,fj->fi", JTw, rhs)
        self.linpoint[mask] = x0

    # ---- gbp/gbp_ba.py:20-34
    def generate_priors_var(self, weaker_factor=100.0):
        fmax = self.factor_lam.reshape(self.F, -1).max(axis=1)
        cmax = np.zeros(self.C); np.maximum.at(cmax, self.cam, fmax)
        lmax = np.zeros(self.L); np.maximum.at(lmax, self.lmk, fmax)
        self.cam_prior_lam = np.eye(6)[None] * (cmax / weaker_factor ** 2)[:, None, None]
        self.lmk_prior_lam = np.eye(3)[None] * (lmax / weaker_factor ** 2)[:, None, None]
        self.cam_prior_eta = np.einsum("vij,vj->vi", self.cam_prior_lam, self.cam_mu)
        self.lmk_prior_eta = np.einsum("vij,vj->vi", self.lmk_prior_lam, self.lmk_mu)

    # ---- gbp/gbp_ba.py:36-42
    def weaken_priors(self, f):
        self.cam_prior_eta *= f; self.cam_prior_lam *= f
        self.lmk_prior_eta *= f; self.lmk_prior_lam *= f

    # ---- gbp/gbp_ba.py:44-52
    def set_priors_var(self, cam_cov, lmk_cov):
        self.cam_prior_lam = np.linalg.inv(np.asarray(cam_cov, dtype=np.float64))
        self.lmk_prior_lam = np.linalg.inv(np.asarray(lmk_cov, dtype=np.float64))
        self.cam_prior_eta = np.einsum("vij,vj->vi", self.cam_prior_lam, self.cam_mu)
        self.lmk_prior_eta = np.einsum("vij,vj->vi", self.lmk_prior_lam, self.lmk_mu)

    # ---- gbp/gbp.py:56-58, 176-198
    def update_all_beliefs(self):
        ce = self.cam_prior_eta.copy(); cl = self.cam_prior_lam.copy()
        le = self.lmk_prior_eta.copy(); ll = self.lmk_prior_lam.copy()
        # np.add.at accumulates in index order = adj_factors order of the reference
        np.add.at(ce, self.cam, self.msg_cam_eta); np.add.at(cl, self.cam, self.msg_cam_lam)
        np.add.at(le, self.lmk, self.msg_lmk_eta); np.add.at(ll, self.lmk, self.msg_lmk_lam)
        self.cam_eta, self.cam_lam, self.lmk_eta, self.lmk_lam = ce, cl, le, ll
        self.cam_Sigma = np.linalg.inv(cl); self.lmk_Sigma = np.linalg.inv(ll)
        self.cam_mu = np.einsum("vij,vj->vi", self.cam_Sigma, ce)
        self.lmk_mu = np.einsum("vij,vj->vi", self.lmk_Sigma, le)

    def adj_means(self):
        """gbp/gbp.py:72-74, 255-257: inv(belief.lam) @ belief.eta of both adjacent beliefs."""
        cm = np.einsum("vij,vj->vi", np.linalg.inv(self.cam_lam), self.cam_eta)
        lm = np.einsum("vij,vj->vi", np.linalg.inv(self.lmk_lam), self.lmk_eta)
        return np.concatenate([cm[self.cam], lm[self.lmk]], axis=1)

    # ---- gbp/gbp.py:82-84, 296-332
    def robustify_all_factors(self):
        old = self.adaptive_var.copy()
        if self.loss is None:
            self.adaptive_var[:] = self.var0
        else:
            pred = meas_fn(self.linpoint, self.K)
            M = np.linalg.norm(self.z - pred, axis=1) / np.sqrt(self.var0)
            over = M > self.Nstds
            if self.loss == "huber":
                with np.errstate(divide="ignore", invalid="ignore"):
                    v = self.var0 * M ** 2 / (2 * (self.Nstds * M - 0.5 * self.Nstds ** 2))
            elif self.loss == "constant":
                v = M ** 2
            else:  # unknown name: the reference leaves the variance (and the factor) unchanged
                return
            self.adaptive_var = np.where(over, v, self.var0)
            self.robust_flag = over
        s = old / self.adaptive_var
        self.factor_eta *= s[:, None]
        self.factor_lam *= s[:, None, None]

    # ---- gbp/gbp.py:64-80
    def relinearise_factors(self):
        means = self.adj_means()
        relin = (np.linalg.norm(self.linpoint - means, axis=1) > self.beta) & \
                (self.iters_since_relin >= self.min_linear_iters)
        self.compute_factor(relin, means)
        self.iters_since_relin = np.where(relin, 0, self.iters_since_relin + 1)
        self.factor_damping = np.where(relin, 0.0, self.factor_damping)
        return relin

    # ---- gbp/gbp.py:46-54, 334-373
    def compute_all_messages(self, local_relin=True):
        if local_relin:
            self.factor_damping = np.where(self.iters_since_relin == self.num_undamped_iters,
                                           self.eta_damping, self.factor_damping)
            damp = self.factor_damping
        else:
            damp = np.full(self.F, self.eta_damping)
        ef, lf = self.factor_eta, self.factor_lam
        cam, lmk = self.cam, self.lmk
        # -- message to the camera (v = 0): add (belief - message) of the landmark, marginalise it
        e = ef.copy(); l = lf.copy()
        e[:, 6:9] += self.lmk_eta[lmk] - self.msg_lmk_eta
        l[:, 6:9, 6:9] += self.lmk_lam[lmk] - self.msg_lmk_lam
        inv_nono = np.linalg.inv(l[:, 6:9, 6:9])
        lono = l[:, 0:6, 6:9]
        new_c_lam = l[:, 0:6, 0:6] - lono @ inv_nono @ l[:, 6:9, 0:6]
        new_c_eta = e[:, 0:6] - np.einsum("fij,fj->fi", lono @ inv_nono, e[:, 6:9])
        new_c_eta = (1 - damp)[:, None] * new_c_eta + damp[:, None] * self.msg_cam_eta
        # -- message to the landmark (v = 1): add (belief - message) of the camera, marginalise it
        e = ef.copy(); l = lf.copy()
        e[:, 0:6] += self.cam_eta[cam] - self.msg_cam_eta
        l[:, 0:6, 0:6] += self.cam_lam[cam] - self.msg_cam_lam
        inv_nono = np.linalg.inv(l[:, 0:6, 0:6])
        lono = l[:, 6:9, 0:6]
        new_l_lam = l[:, 6:9, 6:9] - lono @ inv_nono @ l[:, 0:6, 6:9]
        new_l_eta = e[:, 6:9] - np.einsum("fij,fj->fi", lono @ inv_nono, e[:, 0:6])
        new_l_eta = (1 - damp)[:, None] * new_l_eta + damp[:, None] * self.msg_lmk_eta
        # both messages replaced only after both are computed (gbp/gbp.py:371-373)
        self.msg_cam_eta, self.msg_cam_lam = new_c_eta, new_c_lam
        self.msg_lmk_eta, self.msg_lmk_lam = new_l_eta, new_l_lam

    # ---- gbp/gbp.py:86-92
    def synchronous_iteration(self, local_relin=True, robustify=False):
        if robustify:
            self.robustify_all_factors()
        if local_relin:
            self.relinearise_factors()
        self.compute_all_messages(local_relin=local_relin)
        self.update_all_beliefs()

    # ---- gbp/gbp.py:251-259
    def compute_residuals(self):
        return meas_fn(self.adj_means(), self.K) - self.z

    # ---- gbp/gbp_ba.py:61-69
    def are(self):
        return float(np.sum(np.linalg.norm(self.compute_residuals(), axis=1)) / self.F)

    # ---- gbp/gbp.py:36-44
    def energy(self):
        r = self.compute_residuals()
        return float(np.sum(0.5 * np.linalg.norm(r, axis=1) ** 2 / self.adaptive_var))

    def n_relinearising(self):
        """ba.py:97-100."""
        return int(np.sum(self.iters_since_relin == 0))


def run_ba_loop(o: BAOracle, n_iters, weaker_factor=50.0, float_impl=False, on_iter=None,
                final_weaker=100.0, n_weak=5):
    """The body of ba.py:75-105 (without the viewer).  Returns per-iteration traces."""
    o.generate_priors_var(weaker_factor)
    o.update_all_beliefs()
    wf = np.log10(final_weaker) / n_weak                 # ba.py:65
    are, en, nrel = [], [], []
    for i in range(n_iters):
        if float_impl and (i + 1) % 2 == 0 and i < n_weak * 2:   # ba.py:86-88
            o.weaken_priors(wf)
        if i == 3 or i == 8:                                      # ba.py:91-93
            o.iters_since_relin[:] = 1
        are.append(o.are()); en.append(o.energy()); nrel.append(o.n_relinearising())
        o.synchronous_iteration(robustify=True, local_relin=True)
        if on_iter is not None:
            on_iter(i, o)
    are.append(o.are()); en.append(o.energy()); nrel.append(o.n_relinearising())
    return np.array(are), np.array(en), np.array(nrel)
