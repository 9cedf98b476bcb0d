"""GPU parity of the resident kernel (gbp_resident.cuh: the whole gbp_ba_iterate(n) of an L2-resident graph in one cooperative
launch, one grid barrier per iteration) against the two-kernel iteration: same summation orders, so the SAME BITS, through
relinearisations, robust reweighting, client resets and client-written tables; and against the reference fixtures."""
import numpy as np
import pytest

from conftest import golden_configs, golden_problem, load_golden, relerr

pytestmark = pytest.mark.gpu

STATE_FIELDS = ("F_CAM_BELIEF", "F_LMK_BELIEF", "F_MSG_CAM", "F_MSG_LMK", "F_LINPOINT", "F_ITERS", "F_FLAGS", "F_ADAPTIVE_VAR",
                "F_CAM_PARTIAL", "F_CAM_MU", "F_LMK_MU")


def _state(g):
    from gbp_b200 import _lib as L
    return {f: g._eng.read(getattr(L, f)).copy() for f in STATE_FIELDS}


def _assert_same_state(a, b, what=""):
    sa, sb = _state(a), _state(b)
    for f in STATE_FIELDS:
        assert np.array_equal(sa[f], sb[f]), (what, f, float(np.max(np.abs(sa[f].astype(float) - sb[f].astype(float)))))


def _ba_schedule(g, n_iters):
    g.generate_priors_var(50.0)
    g.update_all_beliefs()
    g.iterate(3, robustify=True, local_relin=True); g.reset_iters_since_relin(1)
    g.iterate(5, robustify=True, local_relin=True); g.reset_iters_since_relin(1)
    g.iterate(n_iters - 8, robustify=True, local_relin=True)


@pytest.mark.parametrize("name", ["fr1desk", "fr1desk_vsmall", "fr1desk_vsmall_huber", "fr1desk_vsmall_constant"])
def test_resident_kernel_equals_two_kernel_iteration(name):
    """ba.py's schedule (resets at 3 and 8) for the whole run of the fixture: resident kernel (kernel_variant 0 on a small
    graph) vs the two-kernel iteration (kernel_variant 1): every table bit-identical; final means within 1e-4 of the reference."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden(name)
    n_iters = int(G["n_iters"])
    a = create_ba_graph(golden_problem(G), golden_configs(G), kernel_variant=1)
    b = create_ba_graph(golden_problem(G), golden_configs(G), kernel_variant=0)
    assert a._eng.resident_warps == 0 and b._eng.resident_warps >= 1
    l0 = b._eng.launch_count()
    _ba_schedule(a, n_iters)
    _ba_schedule(b, n_iters)
    assert b._eng.launch_count() - l0 < 40            # three iterate() calls = three resident launches (+ belief kernels, priors)
    _assert_same_state(a, b, name)
    key = f"s{int(G['checkpoints'].max())}"
    assert relerr(b.get_means(), np.concatenate([G[f"{key}_cam_mu"].ravel(), G[f"{key}_lmk_mu"].ravel()])) < 1e-4
    assert abs(b.are() - G["are"][n_iters]) < 1e-6 * G["are"][n_iters]
    a.close(); b.close()


def test_resident_kernel_float_implementation_schedule():
    """--float_implementation: priors weakened between iterate() calls (ba.py:86-88): the first iteration of a resident call
    reads the stored beliefs (old priors), like the reference does."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk_vsmall_float")
    wf = np.log10(100.0) / 5
    gs = [create_ba_graph(golden_problem(G), golden_configs(G), kernel_variant=v) for v in (1, 0)]
    for g in gs:
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        done = 0
        for i in range(30):
            if (i + 1) % 2 == 0 and i < 10:
                g.weaken_priors(wf)
            if i in (3, 8):
                g.reset_iters_since_relin(1)
            if i < 10:
                g.iterate(1, robustify=True, local_relin=True); done += 1
            elif i == 10:
                g.iterate(20, robustify=True, local_relin=True); done += 20
        assert done == 30
    _assert_same_state(gs[0], gs[1])
    assert relerr(gs[1].get_means(), np.concatenate([G["s29_cam_mu"].ravel(), G["s29_lmk_mu"].ravel()])) < 1e-4
    for g in gs:
        g.close()


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_resident_kernel_tiles_per_cta_do_not_change_results(warps):
    from gbp_b200 import _lib as L
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk_vsmall_huber")
    a = create_ba_graph(golden_problem(G), golden_configs(G))
    b = create_ba_graph(golden_problem(G), golden_configs(G))
    b._eng.tune(L.TUNE_RESIDENT_WARPS, warps)
    for g in (a, b):
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        g.iterate(17, robustify=True, local_relin=True)
        g.iterate(2, robustify=True, local_relin=True)
        g.iterate(6, robustify=False, local_relin=False)
    _assert_same_state(a, b, warps)
    a.close(); b.close()


def test_resident_kernel_respects_client_written_tables_and_switch():
    """A belief / message table written by the client is what the next sweep reads (iteration 0 of a resident call reads the
    stored beliefs); gbp_ba_tune(GBP_TUNE_RESIDENT, 0) falls back to the two-kernel iteration on the same handle."""
    from gbp_b200 import _lib as L
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk_vsmall")
    a = create_ba_graph(golden_problem(G), golden_configs(G))
    b = create_ba_graph(golden_problem(G), golden_configs(G))
    a._eng.tune(L.TUNE_RESIDENT, 0)
    rng = np.random.default_rng(0)
    for g in (a, b):
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        g.iterate(4, robustify=True, local_relin=True)
    lb = a._eng.read(L.F_LMK_BELIEF).copy()
    lb[:, 9:] += 1e-3 * rng.standard_normal(lb[:, 9:].shape)        # move the means: relinearisation decisions change
    mc = a._eng.read(L.F_MSG_CAM).copy() * 0.5
    for g in (a, b):
        g._eng.write(L.F_LMK_BELIEF, lb)
        g._eng.write(L.F_MSG_CAM, mc)
        g.reset_iters_since_relin(8)
        g.iterate(5, robustify=True, local_relin=True)
    _assert_same_state(a, b)
    l0 = a._eng.launch_count(); a.iterate(10, robustify=True, local_relin=True); two = a._eng.launch_count() - l0
    l0 = b._eng.launch_count(); b.iterate(10, robustify=True, local_relin=True); one = b._eng.launch_count() - l0
    assert two == 20 and one == 2
    _assert_same_state(a, b)
    a.close(); b.close()


def test_resident_kernel_ragged_and_isolated():
    """Ragged tiles, a keyframe with several tiles, landmarks of degree > 8 (several chunks), an isolated landmark and an
    isolated keyframe (their beliefs stay at the prior), odd and even iteration counts (final parity of the double buffers)."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.balio import BALProblem
    from gbp_b200.synthetic import make_synthetic
    p = make_synthetic(12, 150, 11, seed=4)                # every landmark seen by 11 of 12 keyframes: 2 chunks each
    cam = np.vstack([p.cam_means, p.cam_means[:1] + 0.01])   # + an isolated keyframe
    lmk = np.vstack([p.lmk_means, [[0.1, 0.2, 0.3]]])        # + an isolated landmark
    prob = BALProblem(p.cam_id, p.lmk_id, p.z, cam, lmk, p.K4)
    cfg = dict(gauss_noise_std=2, loss="huber", Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
    for n in (2, 3, 16):
        a = create_ba_graph(prob, cfg, kernel_variant=1)
        b = create_ba_graph(prob, cfg)
        assert b._eng.resident_warps >= 1
        for g in (a, b):
            g.generate_priors_var(50.0)
            g.update_all_beliefs()
            g.iterate(n, robustify=True, local_relin=True)
        _assert_same_state(a, b, n)
        assert np.isfinite(b.get_means()).all()
        a.close(); b.close()
