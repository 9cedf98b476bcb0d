"""CPU checks of the arithmetic the CUDA kernels run.

`tests/host_harness/harness.cpp` compiles the product's device headers (`gbp_math.cuh`, `gbp_edge.cuh`: every
function is `__host__ __device__`) with g++ and drives them with plain host loops -- test infrastructure only,
never loaded by the package.  So without a GPU this suite already follows the per-edge step (robustify ->
relinearise -> both messages, `edge_sweep`) through complete `ba.py` runs against the fixtures generated from the
unmodified reference.  The kernels' plumbing (tiles, bulk copies, reductions) is what `-m gpu` covers.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_configs, load_golden, relerr
from oracle import gbp_oracle as O

HARNESS = os.path.join(ROOT, "tests", "host_harness")
SO = os.path.join(HARNESS, "_build", "libgbp_math_host.so")
DEPS = [os.path.join(HARNESS, "harness.cpp")] + [os.path.join(ROOT, "gbp_b200", "csrc", h) for h in ("gbp_math.cuh", "gbp_edge.cuh")]
K4 = np.array([517.306408, 516.469215, 318.64304, 255.313989])
LOSS = {None: 0, "huber": 1, "constant": 2}
_IU6, _IU3 = np.triu_indices(6), np.triu_indices(3)


@pytest.fixture(scope="module")
def hh():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", DEPS[0], "-o", SO], check=True)
    lib = C.CDLL(SO)
    vp = C.c_void_p
    lib.hh_linearise.argtypes = [vp, C.c_long, vp, vp, vp]
    lib.hh_messages.argtypes = [vp, vp, C.c_long, vp, C.c_double, vp, vp, vp, vp, vp, vp, vp]
    lib.hh_message_downdated.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_long, vp, vp]
    lib.hh_solve6.argtypes = lib.hh_solve3.argtypes = [vp, vp, C.c_long, vp]
    lib.hh_robust_variance.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int)]
    lib.hh_robust_variance.restype = C.c_double
    lib.hs_create.argtypes = [C.c_double] * 4 + [C.c_int] * 5 + [C.c_long] + [vp] * 6
    lib.hs_create.restype = vp
    lib.hs_destroy.argtypes = [vp]
    lib.hs_generate_priors.argtypes = lib.hs_scale_priors.argtypes = [vp, C.c_double]
    lib.hs_update_beliefs.argtypes = [vp]
    lib.hs_iterate.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.hs_fill_iters.argtypes = lib.hs_sweep.argtypes = [vp, C.c_int]
    lib.hs_metrics.argtypes = [vp, vp]
    lib.hs_read.argtypes = lib.hs_read_int.argtypes = [vp, C.c_int, vp]
    lib.hs_set_factored.argtypes = [vp, C.c_int]
    lib.hh_factor_roundtrip.argtypes = [vp, C.c_long, vp, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _sym(packed, n):
    iu = _IU6 if n == 6 else _IU3
    out = np.zeros(packed.shape[:-1] + (n, n))
    out[..., iu[0], iu[1]] = packed
    out[..., iu[1], iu[0]] = packed
    return out


class HostSweep:
    """ba.py-shaped driver of the harness' whole-graph loops (factor order = camera-major like the reference)."""

    def __init__(self, lib, G, cfg, factored=False):
        self.lib = lib
        cam_id = np.asarray(G["in_cam_id"], dtype=np.int64)
        order = np.argsort(cam_id, kind="stable")
        self.cam = np.ascontiguousarray(cam_id[order], dtype=np.int32)
        self.lmk = np.ascontiguousarray(np.asarray(G["in_lmk_id"])[order], dtype=np.int32)
        z = np.ascontiguousarray(np.asarray(G["in_z"], dtype=np.float64)[order])
        cam0 = np.ascontiguousarray(G["in_cam0"], dtype=np.float64)
        lmk0 = np.ascontiguousarray(G["in_lmk0"], dtype=np.float64)
        k4 = np.ascontiguousarray(G["in_K"], dtype=np.float64).reshape(4)      # fx fy cx cy
        self.C, self.L, self.F = len(cam0), len(lmk0), len(self.cam)
        self.eta_damping = float(cfg["eta_damping"])
        self.h = lib.hs_create(float(cfg["gauss_noise_std"]), self.eta_damping, float(cfg["beta"]), float(cfg.get("Nstds", 3.0)),
                               int(cfg["num_undamped_iters"]), int(cfg["min_linear_iters"]), LOSS[cfg.get("loss")],
                               self.C, self.L, self.F, _p(self.cam), _p(self.lmk), _p(z), _p(cam0), _p(lmk0), _p(k4))
        if factored:
            lib.hs_set_factored(self.h, 1)

    def close(self):
        self.lib.hs_destroy(self.h)

    def read(self, field, rows, width):
        out = np.empty((rows, width))
        self.lib.hs_read(self.h, field, _p(out))
        return out

    def read_int(self, field):
        out = np.empty(self.F, dtype=np.int32)
        self.lib.hs_read_int(self.h, field, _p(out))
        return out

    def metrics(self):
        out = np.zeros(3)
        self.lib.hs_metrics(self.h, _p(out))
        return out[0] / self.F, out[1], int(round(out[2]))

    def run(self, n_iters, weaker, float_impl=False, on_iter=None):
        """The body of ba.py:75-105 without the viewer."""
        lib, h = self.lib, self.h
        lib.hs_generate_priors(h, float(weaker))
        lib.hs_update_beliefs(h)
        wf = np.log10(100.0) / 5
        tr = []
        for i in range(n_iters):
            if float_impl and (i + 1) % 2 == 0 and i < 10:
                lib.hs_scale_priors(h, wf)
            if i == 3 or i == 8:
                lib.hs_fill_iters(h, 1)
            tr.append(self.metrics())
            lib.hs_iterate(h, 1, 1, 1)
            if on_iter:
                on_iter(i)
        tr.append(self.metrics())
        a = np.array(tr)
        return a[:, 0], a[:, 1], a[:, 2].astype(np.int64)


# ---------------------------------------------------------------------------------------------- per-function checks
def test_linearise_matches_reference_model(hh):
    """meas_fn / jac_fn (gbp/factors/reprojection.py:12-44) incl. SURVEY 8(c)'s known answer for factor 0."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(4096, 9))
    x[:, 2] += 3.0
    J, h = np.empty((len(x), 18)), np.empty((len(x), 2))
    hh.hh_linearise(_p(x), len(x), _p(K4), _p(J), _p(h))
    Ko = O.K_matrix(K4)
    ho, Jo = O.meas_fn(x, Ko), O.jac_fn(x, Ko).reshape(-1, 18)
    assert np.max(np.abs(h - ho) / (1 + np.abs(ho))) < 1e-12
    assert np.max(np.abs(J - Jo) / (1 + np.abs(Jo))) < 1e-11
    G = load_golden("fr1desk_vsmall")
    x0 = np.ascontiguousarray(G["init_linpoint"][:1])
    hh.hh_linearise(_p(x0), 1, _p(K4), _p(J), _p(h))
    from test_oracle_golden import KNOWN_J0
    np.testing.assert_allclose(J[0].reshape(2, 9), KNOWN_J0, atol=2e-9)
    np.testing.assert_allclose(h[0], [144.137616243, 327.4150474715], atol=1e-9)


def test_linearise_nan_at_zero_rotation_like_the_reference(hh):
    """dR_wx_dw divides by w.w (utils/derivatives.py:44): NaN at w = 0, no silent fix."""
    x = np.array([[0.1, 0.2, 3.0, 0.0, 0.0, 0.0, 0.3, -0.2, 0.5]])
    J, h = np.empty((1, 18)), np.empty((1, 2))
    hh.hh_linearise(_p(x), 1, _p(K4), _p(J), _p(h))
    assert np.all(np.isfinite(h)) and np.any(np.isnan(J[0, 3:6]))
    with np.errstate(all="ignore"):
        assert np.any(np.isnan(O.jac_fn(x, O.K_matrix(K4))[0, :, 3:6]))


def test_low_rank_messages_equal_the_schur_complement_form(hh):
    """message<> (Woodbury form) against the oracle's explicit 9x9 factor + Schur complements
    (gbp/gbp.py:334-373) on a live state: beliefs and old messages after two sweeps, damping on for half the edges."""
    G = load_golden("fr1desk_vsmall")
    cfg = golden_configs(G)
    o = O.BAOracle(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"], cfg)
    o.generate_priors_var(cfg["prior_std_weaker_factor"])
    o.update_all_beliefs()
    for _ in range(2):
        o.synchronous_iteration(robustify=True, local_relin=True)
    F = o.F
    damp = np.where(np.arange(F) % 2 == 0, 0.4, 0.0)
    bel_c = np.concatenate([o.cam_eta, o.cam_lam[:, _IU6[0], _IU6[1]]], axis=1)[o.cam]
    bel_l = np.concatenate([o.lmk_eta, o.lmk_lam[:, _IU3[0], _IU3[1]]], axis=1)[o.lmk]
    msg_c = np.concatenate([o.msg_cam_eta, o.msg_cam_lam[:, _IU6[0], _IU6[1]]], axis=1)
    msg_l = np.concatenate([o.msg_lmk_eta, o.msg_lmk_lam[:, _IU3[0], _IU3[1]]], axis=1)
    out_c, out_l = np.empty((F, 27)), np.empty((F, 9))
    x0, z = np.ascontiguousarray(o.linpoint), np.ascontiguousarray(o.z)
    hh.hh_messages(_p(x0), _p(z), F, _p(o.K4), o.var0, _p(damp), _p(np.ascontiguousarray(bel_c)), _p(np.ascontiguousarray(bel_l)),
                   _p(msg_c), _p(msg_l), _p(out_c), _p(out_l))
    o.factor_damping = damp.copy()
    o.iters_since_relin[:] = 1           # not num_undamped_iters: compute_all_messages keeps the damping we set
    o.compute_all_messages(local_relin=True)
    assert relerr(out_c[:, :6], o.msg_cam_eta) < 1e-9 and relerr(_sym(out_c[:, 6:], 6), o.msg_cam_lam) < 1e-9
    assert relerr(out_l[:, :3], o.msg_lmk_eta) < 1e-9 and relerr(_sym(out_l[:, 3:], 3), o.msg_lmk_lam) < 1e-9
    # per-message check as well (a global max can hide small messages)
    num = np.abs(_sym(out_l[:, 3:], 3) - o.msg_lmk_lam).reshape(F, -1).max(axis=1)
    den = np.abs(o.msg_lmk_lam).reshape(F, -1).max(axis=1)
    assert np.max(num / den) < 1e-7


def test_downdated_landmark_message_equals_the_explicit_cavity(hh):
    """message_downdated (shared Cholesky factor of the keyframe belief, rank-2 down-date by the old message: what the streaming
    kernel runs) against message<3, 6> on the explicitly formed cavity, for keyframes of many edges (the old message is a small
    share of the belief), of few edges (a large share) and with a zero old message (first sweep)."""
    rng = np.random.default_rng(11)
    n = 3000
    J = rng.normal(size=(n, 18)) * rng.uniform(0.1, 200.0, size=(n, 1))
    b = rng.normal(size=(n, 2)) * 50
    var = rng.uniform(0.5, 40.0, size=n)
    W0 = rng.normal(size=(n, 2, 6)) * rng.uniform(0.05, 30.0, size=(n, 1, 1))
    W0[:100] = 0.0                                                   # zero old message
    share = np.concatenate([np.full(100, 1.0), 10.0 ** rng.uniform(-5, -0.3, size=n - 100)])     # information of the others / old message
    lam_b = np.zeros((n, 6, 6))
    for i in range(n):
        R = rng.normal(size=(6, 6))
        others = R @ R.T + 6 * np.eye(6)
        old = W0[i].T @ W0[i]
        scale = (np.trace(old) / np.trace(others)) / share[i] if np.trace(old) > 0 else 1.0
        lam_b[i] = old + others * max(scale, 1e-12) if np.trace(old) > 0 else others
    iu = np.triu_indices(6)
    lam_p = np.ascontiguousarray(lam_b[:, iu[0], iu[1]])
    e = rng.normal(size=(n, 6)) * 100
    damping = np.where(rng.uniform(size=n) < 0.5, 0.4, 0.0)
    old_eta = rng.normal(size=(n, 3))
    out_a, out_b = np.zeros((n, 9)), np.zeros((n, 9))
    hh.hh_message_downdated(_p(J), _p(b), _p(var), _p(lam_p), _p(np.ascontiguousarray(W0.reshape(n, 12))), _p(e), _p(damping), _p(old_eta),
                            n, _p(out_a), _p(out_b))
    assert np.isfinite(out_a).all() and np.isfinite(out_b).all()
    # cond(M) ~ 1 / (1 - share of this edge): both forms lose those digits; compare at that scale
    tol = 1e-11 / np.minimum(share, 1.0)
    err = np.max(np.abs(out_a - out_b), axis=1) / np.max(np.abs(out_a), axis=1)
    assert np.all(err < tol), (float(err.max()), int(np.argmax(err / tol)))
    assert err[:100].max() < 1e-13 and np.median(err) < 1e-13


@pytest.mark.parametrize("n", [3, 6])
def test_spd_solve(hh, n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(500, n, n))
    A = A @ np.swapaxes(A, 1, 2) + 0.1 * np.eye(n)
    r = rng.normal(size=(500, n))
    iu = np.triu_indices(n)
    P = np.ascontiguousarray(A[:, iu[0], iu[1]])
    x = np.empty((500, n))
    (hh.hh_solve6 if n == 6 else hh.hh_solve3)(_p(P), _p(r), 500, _p(x))
    xo = np.linalg.solve(A, r[..., None])[..., 0]
    assert np.max(np.abs(x - xo) / np.max(np.abs(xo), axis=1, keepdims=True)) < 1e-9


@pytest.mark.parametrize("loss", ["huber", "constant"])
def test_robust_variance(hh, loss):
    """Factor.robustify_loss (gbp/gbp.py:296-332): adaptive variance and robust_flag."""
    var0, nstds = 4.0, 3.0
    for r in ([0.5, 0.5], [5.0, -3.0], [6.0, 0.0], [40.0, 9.0], [0.0, 6.0000001]):
        M = np.hypot(*r) / np.sqrt(var0)
        flag = C.c_int()
        v = hh.hh_robust_variance(LOSS[loss], var0, nstds, r[0], r[1], C.byref(flag))
        if M > nstds:
            want = var0 * M ** 2 / (2 * (nstds * M - 0.5 * nstds ** 2)) if loss == "huber" else M ** 2
            assert flag.value == 1 and abs(v - want) < 1e-12 * want
        else:
            assert flag.value == 0 and v == var0


def test_factored_message_round_trip(hh):
    """A keyframe message written by a client in full form survives the factored layout: rank-2 PSD matrices exactly
    (to rounding), the zero message, rank 1; W^T W of a random W is reproduced although W itself is not unique."""
    rng = np.random.default_rng(5)
    W = rng.normal(size=(300, 2, 6)) * rng.uniform(0.1, 300, size=(300, 1, 1))
    W[0] = 0.0
    W[1, 1] = 0.0
    lam = np.einsum("nki,nkj->nij", W, W)
    packed = np.ascontiguousarray(lam[:, _IU6[0], _IU6[1]])
    Wout, back = np.empty((300, 12)), np.empty((300, 21))
    hh.hh_factor_roundtrip(_p(packed), 300, _p(Wout), _p(back))
    scale = np.maximum(np.abs(packed).max(axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(back - packed) / scale) < 1e-12
    assert np.all(Wout[0] == 0.0) and np.all(np.isfinite(Wout))


# ---------------------------------------------------------------------------------------------- whole trajectories
def _check_state(s, G, key, tol):
    fs = G["fsample"]
    cb, lb = s.read(0, s.C, 33), s.read(1, s.L, 12)
    mc, ml = s.read(4, s.F, 27)[fs], s.read(5, s.F, 9)[fs]
    worst = {
        "cam_mu": relerr(cb[:, 27:].ravel(), G[f"{key}_cam_mu"]), "lmk_mu": relerr(lb[:, 9:].ravel(), G[f"{key}_lmk_mu"]),
        "cam_eta": relerr(cb[:, :6].ravel(), G[f"{key}_cam_eta"]), "lmk_eta": relerr(lb[:, :3].ravel(), G[f"{key}_lmk_eta"]),
        "cam_lam": relerr(_sym(cb[:, 6:27], 6).ravel(), G[f"{key}_cam_lam"]), "lmk_lam": relerr(_sym(lb[:, 3:9], 3).ravel(), G[f"{key}_lmk_lam"]),
        "msg_cam_eta": relerr(mc[:, :6], G[f"{key}_msg_cam_eta"]), "msg_cam_lam": relerr(_sym(mc[:, 6:], 6), G[f"{key}_msg_cam_lam"]),
        "msg_lmk_eta": relerr(ml[:, :3], G[f"{key}_msg_lmk_eta"]), "msg_lmk_lam": relerr(_sym(ml[:, 3:], 3), G[f"{key}_msg_lmk_lam"]),
        "linpoint": relerr(s.read(6, s.F, 9)[fs], G[f"{key}_linpoint"]),
    }
    assert np.array_equal(s.read_int(7), G[f"{key}_iters_since_relin"]), key
    assert np.array_equal(np.where(s.read_int(8) & 1, s.eta_damping, 0.0), G[f"{key}_eta_damping"]), key
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, (key, bad)


@pytest.mark.parametrize("factored", [False, True], ids=["full", "factored"])
@pytest.mark.parametrize("name", ["fr1desk_vsmall", "fr1desk_vsmall_huber", "fr1desk_vsmall_constant", "fr1desk_vsmall_float",
                                  "fr2robot2", "fr1desk_small", "fr1xyz_av"])      # the last three: the reference's other data files
def test_edge_sweep_trajectory_against_reference_fixture(hh, name, factored):
    """Every checkpoint of the reference run (all loss modes, --float_implementation): beliefs, sampled messages and
    linearisation points, every factor's iters_since_relin and damping flag, the ARE / energy / relinearisation traces.
    `factored` = the compressed keyframe-message layout (eta | W with Lambda = W^T W: the streaming build, kernel_variant 2)."""
    G = load_golden(name)
    cfg = golden_configs(G)
    s = HostSweep(hh, G, cfg, factored)
    cks = set(G["checkpoints"].tolist())
    float_impl = bool(G["float_impl"])

    def on_iter(i):
        if i in cks:
            _check_state(s, G, f"s{i}", 1e-7 if i <= 2 else (1e-4 if float_impl else 1e-5))

    are, en, nrel = s.run(int(G["n_iters"]), cfg["prior_std_weaker_factor"], float_impl, on_iter)
    assert np.array_equal(nrel, G["n_relin"])
    # --float_implementation runs with 100x weaker priors: the run is ill-conditioned and one outer iteration (29) sits
    # on an energy spike (4.4e7 between neighbours of 1e5) where rounding differences show up at 1e-3 in the TRACE
    # before contracting again (2.7e-8 one iteration later); states at the checkpoints are held to 1e-4 above
    tol = 1e-3 if float_impl else 1e-6
    assert relerr(are, G["are"]) < tol and relerr(en, G["energy"]) < tol
    if cfg.get("loss") is not None:
        last = int(G["checkpoints"].max())
        assert relerr(s.read(9, s.F, 1)[:, 0], G[f"s{last}_adaptive_var"]) < 1e-5
    s.close()


@pytest.mark.parametrize("factored", [False, True], ids=["full", "factored"])
def test_edge_sweep_fr1desk_200_iterations(hh, factored):
    """BASELINE config 3 on the host build of the device arithmetic: converged means AND precisions within the
    north-star tolerance (1e-4 relative) of the reference, identical relinearisation counts at all 201 reads."""
    G = load_golden("fr1desk")
    cfg = golden_configs(G)
    s = HostSweep(hh, G, cfg, factored)
    are, en, nrel = s.run(200, cfg["prior_std_weaker_factor"])
    assert np.array_equal(nrel, G["n_relin"])
    assert relerr(are, G["are"]) < 1e-6 and relerr(en, G["energy"]) < 1e-6
    cb, lb = s.read(0, s.C, 33), s.read(1, s.L, 12)
    assert relerr(cb[:, 27:].ravel(), G["s199_cam_mu"]) < 1e-4 and relerr(lb[:, 9:].ravel(), G["s199_lmk_mu"]) < 1e-4
    assert relerr(_sym(cb[:, 6:27], 6).ravel(), G["s199_cam_lam"]) < 1e-4 and relerr(_sym(lb[:, 3:9], 3).ravel(), G["s199_lmk_lam"]) < 1e-4
    s.close()


def test_edge_sweep_synthetic_small(hh):
    """Down-scaled instance of the synthetic generator (the pin for BASELINE configs 4-5)."""
    G = load_golden("synth_small")
    cfg = golden_configs(G)
    s = HostSweep(hh, G, cfg)
    cks = set(G["checkpoints"].tolist())
    worst = []

    def on_iter(i):
        if i in cks:
            cb, lb = s.read(0, s.C, 33), s.read(1, s.L, 12)
            worst.append(max(relerr(cb[:, 27:].ravel(), G[f"s{i}_cam_mu"]), relerr(lb[:, 9:].ravel(), G[f"s{i}_lmk_mu"]),
                             relerr(_sym(cb[:, 6:27], 6).ravel(), G[f"s{i}_cam_lam"]), relerr(_sym(lb[:, 3:9], 3).ravel(), G[f"s{i}_lmk_lam"])))
            assert np.array_equal(s.read_int(7), G[f"s{i}_iters_since_relin"])

    are, en, nrel = s.run(int(G["n_iters"]), cfg["prior_std_weaker_factor"], on_iter=on_iter)
    assert np.array_equal(nrel, G["n_relin"])
    assert relerr(are, G["are"]) < 1e-6 and relerr(en, G["energy"]) < 1e-6 and max(worst) < 1e-5
    s.close()


@pytest.mark.parametrize("factored", [False, True], ids=["full", "factored"])
def test_staged_calls_equal_one_sweep(hh, factored):
    """robustify_all_factors / relinearise_factors / compute_all_messages / update_all_beliefs one by one
    (gbp/gbp.py:82-92, the stage bits of gbp_ba_sweep_local) leave the same state as the fused per-edge pass."""
    G = load_golden("fr1desk_vsmall_huber")
    cfg = golden_configs(G)
    a, b = HostSweep(hh, G, cfg, factored), HostSweep(hh, G, cfg, factored)
    for s in (a, b):
        hh.hs_generate_priors(s.h, 50.0)
        hh.hs_update_beliefs(s.h)
    ST_ROBUSTIFY, ST_RELIN, ST_MESSAGES, ST_BELIEFS, ST_LOCAL_DAMPING = 1, 2, 4, 8, 16
    for i in range(20):
        if i == 3:
            hh.hs_fill_iters(a.h, 7); hh.hs_fill_iters(b.h, 7)          # make relinearisation fire early
        hh.hs_iterate(a.h, 1, 1, 1)
        for st in (ST_ROBUSTIFY, ST_RELIN, ST_MESSAGES | ST_LOCAL_DAMPING, ST_BELIEFS):
            hh.hs_sweep(b.h, st)
    for field, rows, w in ((0, a.C, 33), (1, a.L, 12), (4, a.F, 27), (5, a.F, 9), (6, a.F, 9), (9, a.F, 1)):
        assert relerr(b.read(field, rows, w), a.read(field, rows, w)) < 1e-12, field
    assert np.array_equal(a.read_int(7), b.read_int(7)) and np.array_equal(a.read_int(8), b.read_int(8))
    assert (a.read_int(7) < 17).any()
    a.close(); b.close()
