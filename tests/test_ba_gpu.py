"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
inputs and against the fixtures generated from the unmodified reference."""
import numpy as np
import pytest

from conftest import golden_configs, golden_problem, load_golden, relerr, relerr_rows

pytestmark = pytest.mark.gpu

# North-star tolerance (BASELINE.json): converged beliefs within 1e-4 relative of the reference.
TOL_CONVERGED = 1e-4
# Early sweeps (before rounding differences have been amplified by relinearisation decisions)
TOL_EARLY = 1e-7


def _unpack(rows, n):
    from gbp_b200.engine import unpack_sym
    return rows[:, :n], unpack_sym(rows[:, n:n + n * (n + 1) // 2], n), rows[:, -n:]


def _run_loop(graph, G, n_iters, on_iter=None):
    """The body of ba.py:75-105 against our BAFactorGraph (per-iteration metrics like ba.py:95-101)."""
    cfg = golden_configs(G)
    float_impl = bool(G["float_impl"])
    graph.generate_priors_var(weaker_factor=cfg["prior_std_weaker_factor"])
    graph.update_all_beliefs()
    wf = np.log10(100.0) / 5
    are, en, nrel = [], [], []
    for i in range(n_iters):
        if float_impl and (i + 1) % 2 == 0 and i < 10:
            graph.weaken_priors(wf)
        if i == 3 or i == 8:
            for factor in graph.factors:
                factor.iters_since_relin = 1
        a, e, n = graph.metrics()
        are.append(a); en.append(e); nrel.append(n)
        graph.synchronous_iteration(robustify=True, local_relin=True)
        if on_iter:
            on_iter(i)
    a, e, n = graph.metrics()
    are.append(a); en.append(e); nrel.append(n)
    return np.array(are), np.array(en), np.array(nrel)


def _check_snapshot(graph, G, key, tol):
    from gbp_b200 import _lib as L
    fs = G["fsample"]
    ce, cl, cm = _unpack(graph._eng.read(L.F_CAM_BELIEF), 6)
    le, ll, lm = _unpack(graph._eng.read(L.F_LMK_BELIEF), 3)
    worst = {
        "cam_mu": relerr(cm.ravel(), G[f"{key}_cam_mu"]), "lmk_mu": relerr(lm.ravel(), G[f"{key}_lmk_mu"]),
        "cam_eta": relerr(ce.ravel(), G[f"{key}_cam_eta"]), "lmk_eta": relerr(le.ravel(), G[f"{key}_lmk_eta"]),
        "cam_lam": relerr(cl.ravel(), G[f"{key}_cam_lam"]), "lmk_lam": relerr(ll.ravel(), G[f"{key}_lmk_lam"]),
    }
    mce, mcl, _ = _unpack(graph._eng.read(L.F_MSG_CAM)[fs], 6)
    mle, mll, _ = _unpack(graph._eng.read(L.F_MSG_LMK)[fs], 3)
    worst["msg_cam_eta"] = relerr(mce, G[f"{key}_msg_cam_eta"]); worst["msg_cam_lam"] = relerr(mcl, G[f"{key}_msg_cam_lam"])
    worst["msg_lmk_eta"] = relerr(mle, G[f"{key}_msg_lmk_eta"]); worst["msg_lmk_lam"] = relerr(mll, G[f"{key}_msg_lmk_lam"])
    worst["linpoint"] = relerr(graph._eng.read(L.F_LINPOINT)[fs], G[f"{key}_linpoint"])
    assert np.array_equal(graph._eng.read(L.F_ITERS)[:, 0], G[f"{key}_iters_since_relin"]), key
    damp = np.where(graph._eng.read(L.F_FLAGS)[:, 0] & 1, graph.eta_damping, 0.0)
    assert np.array_equal(damp, G[f"{key}_eta_damping"]), key
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, (key, bad)
    return worst


def test_reprojection_model_matches_oracle():
    """meas_fn / jac_fn (gbp/factors/reprojection.py:12-44) evaluated by the kernels."""
    from gbp_b200.engine import reprojection_eval
    from oracle import gbp_oracle as O
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(4096, 9))
    x[:, 2] += 3.0           # keep the point in front of the camera
    K4 = np.array([517.306408, 516.469215, 318.64304, 255.313989])
    h, J = reprojection_eval(x, K4)
    Ko = O.K_matrix(K4)
    ho, Jo = O.meas_fn(x, Ko), O.jac_fn(x, Ko)
    assert np.max(np.abs(h - ho) / (1 + np.abs(ho))) < 1e-12
    assert np.max(np.abs(J - Jo) / (1 + np.abs(Jo))) < 1e-11


def test_reprojection_special_branches_on_the_device():
    """The branches of so3exp / dR_wx_dw that random inputs never reach, evaluated by the DEVICE code (sincos / rsqrt intrinsics, not
    the host build's libm): |w| < 3 eps -> R = I (utils/lie_algebra.py:34-35); w = 0 -> the w-block of the Jacobian is NaN exactly
    like the reference (utils/derivatives.py:43-44: 0 / 0); a rotation just above the threshold takes the Rodrigues branch."""
    from gbp_b200.engine import reprojection_eval
    from oracle import gbp_oracle as O
    x = np.zeros((4, 9))
    x[:, 2] = 3.0
    x[:, 6:9] = [0.3, -0.2, 1.0]
    x[1, 3:6] = [1e-16, 0.0, 0.0]                 # |w| < 3 eps, w != 0: identity rotation, finite Jacobian
    x[2, 3:6] = [1e-9, -2e-9, 5e-10]              # just above the threshold
    x[3, 3:6] = [6.0e-16, 2.0e-16, 1.0e-16]       # |w| = 6.4e-16 < 6.66e-16: still the identity branch
    K4 = np.array([517.306408, 516.469215, 318.64304, 255.313989])
    h, J = reprojection_eval(x, K4)
    Ko = O.K_matrix(K4)
    with np.errstate(all="ignore"):
        ho, Jo = O.meas_fn(x, Ko), O.jac_fn(x, Ko)
    assert np.array_equal(np.isnan(J), np.isnan(Jo)) and np.isnan(J[0, :, 3:6]).all() and not np.isnan(J[1:]).any()
    assert not np.isnan(h).any()
    assert np.max(np.abs(h - ho) / (1 + np.abs(ho))) < 1e-12
    ok = ~np.isnan(Jo)
    assert np.max(np.abs(J[ok] - Jo[ok]) / (1 + np.abs(Jo[ok]))) < 1e-9
    assert np.array_equal(h[0], h[1]) and np.array_equal(h[0], h[3])          # all three used R = I


def test_known_answer_factor0():
    """SURVEY section 8(c) known answers for factor 0 of fr1desk_vsmall through the proxy objects."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk_vsmall")
    graph = create_ba_graph(golden_problem(G), golden_configs(G))
    f0 = graph.factors[0]
    assert f0.adj_vIDs == [0, 47] and f0.iters_since_relin == 1 and f0.eta_damping == 0.0
    np.testing.assert_allclose(f0.measurement, [358.3182, 189.9086], atol=1e-10)
    fs = G["fsample"]
    assert fs[0] == 0
    assert relerr(f0.factor.eta, G["init_factor_eta"][0]) < 1e-12
    assert relerr(f0.factor.lam, G["init_factor_lam"][0]) < 1e-12
    assert abs(f0.factor.lam.max() - 90516.99360510931) < 1e-6
    np.testing.assert_allclose(f0.linpoint, G["init_linpoint"][0], rtol=0, atol=0)
    graph.generate_priors_var(weaker_factor=50.0)
    assert abs(graph.cam_nodes[0].prior.lam[0, 0] - 232.48310953175482) < 1e-9
    assert abs(graph.lmk_nodes[0].prior.lam[0, 0] - 33.684116437247276) < 1e-9
    assert graph.lmk_nodes[0].variableID == 10
    graph.update_all_beliefs()
    graph.synchronous_iteration(robustify=True, local_relin=True)
    np.testing.assert_allclose(graph.cam_nodes[0].mu, [0.1573795672, -0.1153594255, 0.3999238657, -0.1253963241,
                                                       0.185235196, 0.0091166874], atol=2e-9)
    np.testing.assert_allclose(graph.lmk_nodes[0].mu, [-0.5037869154, -0.0611332514, 0.6451349878], atol=2e-9)
    np.testing.assert_allclose(f0.messages[1].eta, [27.3858176992, -16.8956565791, 8.9587025115], atol=2e-8)
    graph.close()


@pytest.mark.parametrize("name", ["fr1desk_vsmall", "fr1desk_vsmall_huber", "fr1desk_vsmall_constant", "fr1desk_vsmall_float",
                                  "fr2robot2", "fr1desk_small", "fr1xyz_av"])      # the last three: the reference's other data files
def test_trajectory_against_reference_fixture(name):
    """Every checkpoint of the reference run: beliefs (eta, Lambda, mu), sampled messages and
    linearisation points, and the per-factor relinearisation / damping state, for all loss modes."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden(name)
    graph = create_ba_graph(golden_problem(G), golden_configs(G))
    cks = set(G["checkpoints"].tolist())
    float_impl = bool(G["float_impl"])

    def on_iter(i):
        if i in cks:
            tol = TOL_EARLY if i <= 2 else (TOL_CONVERGED if float_impl else 1e-5)
            _check_snapshot(graph, G, f"s{i}", tol)

    are, en, nrel = _run_loop(graph, G, int(G["n_iters"]), on_iter)
    assert np.array_equal(nrel, G["n_relin"])
    tol = TOL_CONVERGED if float_impl else 1e-6
    assert relerr(are, G["are"]) < tol and relerr(en, G["energy"]) < tol
    if G["cfg_vals"][list(G["cfg_keys"]).index("loss")] != "None":
        from gbp_b200 import _lib as L
        last = int(G["checkpoints"].max())
        assert relerr(graph._eng.read(L.F_ADAPTIVE_VAR)[:, 0], G[f"s{last}_adaptive_var"]) < 1e-5
    graph.close()


def test_fr1desk_200_iterations_converged_beliefs():
    """BASELINE config 3: fr1desk, defaults, 200 synchronous iterations; beliefs (mean AND precision)
    within 1e-4 relative of the reference, same ARE / energy / relinearisation trace."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk")
    graph = create_ba_graph(golden_problem(G), golden_configs(G))
    assert (len(graph.cam_nodes), len(graph.lmk_nodes), len(graph.factors)) == (63, 2869, 13298)
    cks = set(G["checkpoints"].tolist())
    seen = {}

    def on_iter(i):
        if i in cks:
            seen[i] = _check_snapshot(graph, G, f"s{i}", TOL_EARLY if i == 0 else TOL_CONVERGED)

    are, en, nrel = _run_loop(graph, G, 200, on_iter)
    assert sorted(seen) == [0, 15, 16, 99, 199]
    assert abs(are[-1] - 1.656861417438452) < 1e-6 and abs(en[-1] - 7128.308364695148) < 1e-2
    assert relerr(are, G["are"]) < 1e-6 and relerr(en, G["energy"]) < 1e-6
    # relinearisation is a threshold decision (|linpoint - mu| > beta) on a chaotic trajectory: a factor whose distance sits within
    # rounding of beta may flip one iteration earlier / later than in the reference.  Bound both the size and the number of such reads.
    dn = np.abs(nrel - G["n_relin"])
    assert dn.max() <= 2 and int((dn > 0).sum()) <= 10, (dn.max(), int((dn > 0).sum()), np.nonzero(dn)[0][:20])
    # all factors relinearise in sweep 15 (seen by the client at the start of outer iteration 16)
    assert nrel[16] == 13298 and nrel[:16].sum() == 0
    # north-star gate PER VARIABLE (not only over the whole table): every keyframe's / landmark's converged mean and
    # precision block within 1e-4 relative of the reference in its own norm: ||dmu_v|| / ||mu_v||, ||dLam_v||_F / ||Lam_v||_F
    from gbp_b200 import _lib as L
    ce, cl, cm = _unpack(graph._eng.read(L.F_CAM_BELIEF), 6)
    le, ll, lm = _unpack(graph._eng.read(L.F_LMK_BELIEF), 3)
    per_var = {"cam_mu": relerr_rows(cm, G["s199_cam_mu"].reshape(-1, 6)), "lmk_mu": relerr_rows(lm, G["s199_lmk_mu"].reshape(-1, 3)),
               "cam_lam": relerr_rows(cl, G["s199_cam_lam"].reshape(-1, 36)), "lmk_lam": relerr_rows(ll, G["s199_lmk_lam"].reshape(-1, 9)),
               "cam_eta": relerr_rows(ce, G["s199_cam_eta"].reshape(-1, 6)), "lmk_eta": relerr_rows(le, G["s199_lmk_eta"].reshape(-1, 3))}
    assert all(v < TOL_CONVERGED for v in per_var.values()), per_var
    graph.close()


@pytest.mark.parametrize("tile,block", [(32, 0), (64, 0), (128, 0), (32, 100), (128, 64)])
def test_tiling_and_landmark_blocking_do_not_change_results(tile, block):
    """The engine's storage order (tile size, landmark L2 blocks) is invisible in the results."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L
    G = load_golden("fr1desk_vsmall")
    ref = create_ba_graph(golden_problem(G), golden_configs(G), tile_edges=32, lmk_block=0)
    alt = create_ba_graph(golden_problem(G), golden_configs(G), tile_edges=tile, lmk_block=block)
    for g in (ref, alt):
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        g.iterate(20, robustify=True, local_relin=True)
    for f in (L.F_CAM_BELIEF, L.F_LMK_BELIEF, L.F_MSG_CAM, L.F_MSG_LMK, L.F_LINPOINT):
        assert relerr(alt._eng.read(f), ref._eng.read(f)) < 1e-9, f
    assert np.array_equal(alt._eng.read(L.F_ITERS), ref._eng.read(L.F_ITERS))
    assert np.array_equal(alt._eng.read(L.F_ADJ), ref._eng.read(L.F_ADJ))
    ref.close(); alt.close()


def test_kernel_variants_agree():
    """Tile size does not change the results beyond summation order; the streaming build (factored keyframe messages)
    agrees with the full-row build to rounding."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L
    G = load_golden("fr1desk_vsmall_huber")
    gs = [create_ba_graph(golden_problem(G), golden_configs(G), tile_edges=t, kernel_variant=v)
          for t, v in ((64, 1), (64, 0), (128, 1), (32, 1), (64, 2), (32, 2))]
    for g in gs:
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        g.iterate(8); g._eng.fill_iters(8); g.iterate(12, robustify=True)
    for f in (L.F_CAM_BELIEF, L.F_LMK_BELIEF, L.F_MSG_CAM, L.F_MSG_LMK, L.F_LINPOINT, L.F_ADAPTIVE_VAR, L.F_ITERS, L.F_FLAGS):
        assert np.array_equal(gs[0]._eng.read(f), gs[1]._eng.read(f)), f     # automatic choice on a small graph = build 1
        for k in (2, 3, 4, 5):
            assert relerr(gs[k]._eng.read(f), gs[0]._eng.read(f)) < 1e-9, (f, k)
    for g in gs:
        g.close()


def test_synthetic_small_against_oracle():
    """Down-scaled instance of the synthetic generator (configs 4-5): GPU vs oracle, 30 sweeps."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.synthetic import make_synthetic
    from oracle.gbp_oracle import BAOracle, run_ba_loop
    prob = make_synthetic(20, 2000, 10, seed=0)
    cfg = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
               eta_damping=0.4, prior_std_weaker_factor=50.0)
    o = BAOracle(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, cfg)
    are_o, en_o, nrel_o = run_ba_loop(o, 30, 50.0)
    graph = create_ba_graph(prob, cfg, tile_edges=64, lmk_block=512)
    G = {"cfg_keys": np.array(sorted(cfg)), "cfg_vals": np.array([str(cfg[k]) for k in sorted(cfg)]), "float_impl": 0}
    are, en, nrel = _run_loop(graph, G, 30)
    assert np.array_equal(nrel, nrel_o)
    assert relerr(are, are_o) < 1e-6 and relerr(en, en_o) < 1e-6
    mu = graph.get_means()
    assert relerr(mu, np.concatenate([o.cam_mu.ravel(), o.lmk_mu.ravel()])) < 1e-6
    assert en[-1] < 0.1 * en[0]
    graph.close()


def test_staged_calls_equal_fused_iteration():
    """robustify_all_factors / relinearise_factors / compute_all_messages / update_all_beliefs called
    one by one (gbp/gbp.py:82-92) give the same state as synchronous_iteration."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L
    G = load_golden("fr1desk_vsmall_huber")
    a = create_ba_graph(golden_problem(G), golden_configs(G))
    b = create_ba_graph(golden_problem(G), golden_configs(G))
    for g in (a, b):
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
    for i in range(20):
        if i == 3:
            a._eng.fill_iters(7); b._eng.fill_iters(7)      # make relinearisation fire early
        a.synchronous_iteration(robustify=True, local_relin=True)
        b.robustify_all_factors(); b.relinearise_factors(); b.compute_all_messages(local_relin=True); b.update_all_beliefs()
    for f in (L.F_CAM_BELIEF, L.F_LMK_BELIEF, L.F_MSG_CAM, L.F_MSG_LMK, L.F_LINPOINT, L.F_ADAPTIVE_VAR):
        assert relerr(b._eng.read(f), a._eng.read(f)) < 1e-12, f
    assert np.array_equal(a._eng.read(L.F_ITERS), b._eng.read(L.F_ITERS))
    assert (a._eng.read(L.F_ITERS) == 0).any() or (a._eng.read(L.F_ITERS) < 17).any()
    a.close(); b.close()


def test_proxy_api_surface():
    """Attributes the reference clients touch (ba.py:70-105, vis/ba_vis.py:41-43,115)."""
    from gbp_b200.ba import create_ba_graph
    G = load_golden("fr1desk_vsmall")
    graph = create_ba_graph(golden_problem(G), golden_configs(G))
    assert graph.n_var_nodes == 650 and graph.n_factor_nodes == 1801 and graph.n_edges == 3602
    assert len(graph.var_nodes) == 650 and graph.var_nodes[10].variableID == 10 and graph.var_nodes[10].dofs == 3
    assert graph.cam_nodes[3].c_id == 3 and graph.lmk_nodes[5].l_id == 5
    assert graph.factors[0].args[0].shape == (3, 3)
    graph.generate_priors_var(50.0)
    graph.update_all_beliefs()
    for factor in graph.factors:
        factor.iters_since_relin = 5
    graph.factors[7].iters_since_relin = 2
    graph.synchronous_iteration(robustify=True, local_relin=True)
    its = [f.iters_since_relin for f in graph.factors]
    assert its[7] == 3 and its[0] == 6 and set(its) == {3, 6}
    assert graph.factors[0].eta_damping == 0.4 and graph.factors[7].eta_damping == 0.0
    # adjacency lists are consistent with the factors
    v = graph.lmk_nodes[37]
    assert all(f.adj_vIDs[1] == v.variableID for f in v.adj_factors) and len(v.adj_factors) >= 1
    # prior write-through (ndim_posegraph.py:71-72 style assignment)
    node = graph.cam_nodes[1]
    node.prior.lam = np.eye(6) * 3.0
    node.prior.eta = np.arange(6.0)
    graph.update_all_beliefs()
    np.testing.assert_allclose(graph.cam_nodes[1].prior.lam, np.eye(6) * 3.0)
    lam = graph.cam_nodes[1].belief.lam
    assert np.allclose(lam, lam.T) and np.all(np.linalg.eigvalsh(lam) > 0)
    assert np.allclose(graph.cam_nodes[1].Sigma @ lam, np.eye(6), atol=1e-8)
    r = graph.factors[0].compute_residual()
    assert r.shape == (2,) and abs(np.linalg.norm(r) - graph.factors[0].reprojection_err()) < 1e-12
    assert abs(np.mean([np.linalg.norm(x) for x in np.array(graph.compute_residuals()).reshape(-1, 2)]) - graph.are()) < 1e-9
    graph.close()


def test_belief_write_keeps_the_mean_consistent():
    """A client that rebinds node.belief.eta / .lam (or node.mu) on a live graph: every consumer of the reference takes the mean
    as Lambda^-1 eta (gbp/gbp.py:71, 192-193), so after the write-through the device row's mean is Lambda^-1 eta of what was
    written; an assigned mu alone stands only while Lambda is still zero (the initial state)."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L
    G = load_golden("fr1desk_vsmall")
    g = create_ba_graph(golden_problem(G), golden_configs(G))
    n0 = g.cam_nodes[0]
    n0.mu = n0.mu + 0.25                                   # initial state: Lambda = 0, the written mean is kept
    g._flush()
    assert np.allclose(g._eng.read(L.F_CAM_MU)[0], G["in_cam0"][0] + 0.25, rtol=0, atol=0)
    g.reset()
    g.generate_priors_var(50.0); g.update_all_beliefs()
    g.iterate(3, robustify=True, local_relin=True)
    n = g.cam_nodes[1]
    eta, lam, mu = n.belief.eta, n.belief.lam, n.mu
    n.belief.eta = 2.0 * eta
    n.mu = mu + 5.0                                        # no effect where Lambda > 0: the sweep would never see it in the reference either
    g._flush()
    row = g._eng.read(L.F_CAM_BELIEF)[1]
    want = np.linalg.solve(lam, 2.0 * eta)
    assert np.allclose(row[27:], want, rtol=1e-10, atol=1e-12) and np.allclose(row[:6], 2.0 * eta, rtol=0, atol=0)
    assert np.array_equal(g._eng.read(L.F_CAM_MU)[1], row[27:])
    other = g._eng.read(L.F_CAM_BELIEF)[2]
    assert np.allclose(other[27:], np.linalg.solve(g.cam_nodes[2].belief.lam, other[:6]), rtol=1e-10, atol=1e-12)
    g.close()


def test_edge_cases():
    """Empty measurement list, a landmark seen once, a keyframe with no measurements, ragged tiles."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.balio import BALProblem
    from gbp_b200 import _lib as L
    from oracle.gbp_oracle import BAOracle
    cfg = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
               eta_damping=0.4)
    G = load_golden("fr1desk_vsmall")
    P = golden_problem(G)
    # (a) empty graph: creation and metrics work, nothing to sweep
    empty = create_ba_graph(BALProblem(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2)), P.cam_means[:2],
                                       P.lmk_means[:3], P.K4), cfg)
    assert len(empty.factors) == 0 and empty._eng.metrics() == (0.0, 0.0, 0)
    empty.close()
    # (b) 33 measurements of one keyframe (one full 32-tile + a tile of 1) plus an isolated keyframe;
    #     file order deliberately NOT camera-sorted
    sel = np.nonzero(P.cam_id == 2)[0][:33]
    extra = np.nonzero(P.cam_id == 0)[0][:5]
    idx = np.concatenate([sel[:10], extra, sel[10:]])
    used = np.unique(P.lmk_id[idx])
    remap = -np.ones(P.n_points, dtype=np.int32); remap[used] = np.arange(len(used))
    prob = BALProblem(P.cam_id[idx], remap[P.lmk_id[idx]], P.z[idx], P.cam_means[:4], P.lmk_means[used], P.K4)
    g = create_ba_graph(prob, cfg, tile_edges=32)
    o = BAOracle(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, cfg)
    assert np.array_equal(g._eng.read(L.F_FILE_INDEX)[:, 0], o.file_order)
    # keyframes 1 and 3 have no measurement: give every variable an explicit prior instead
    cov = [np.eye(6) * 0.01] * 4 + [np.eye(3) * 0.04] * len(used)
    g.set_priors_var(cov)
    o.set_priors_var(np.stack(cov[:4]), np.stack(cov[4:]))
    g.update_all_beliefs(); o.update_all_beliefs()
    for _ in range(12):
        g.synchronous_iteration(robustify=True, local_relin=True)
        o.synchronous_iteration(robustify=True, local_relin=True)
    assert relerr(g.get_means(), np.concatenate([o.cam_mu.ravel(), o.lmk_mu.ravel()])) < 1e-8
    np.testing.assert_allclose(g.cam_nodes[3].mu, prob.cam_means[3], atol=1e-12)   # untouched keyframe = its prior
    assert abs(g.are() - o.are()) < 1e-8 * o.are()
    g.close()


def test_error_behaviour():
    from gbp_b200 import _lib as L
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.balio import BALProblem
    G = load_golden("fr1desk_vsmall")
    P = golden_problem(G)
    cfg = golden_configs(G)
    bad = BALProblem(P.cam_id.copy(), P.lmk_id.copy(), P.z, P.cam_means, P.lmk_means, P.K4)
    bad.lmk_id[5] = P.n_points + 3
    with pytest.raises(L.GbpError, match="landmark id"):
        create_ba_graph(bad, cfg)
    with pytest.raises(ValueError, match="unknown loss"):
        create_ba_graph(P, dict(cfg, loss="tukey"))
    g = create_ba_graph(P, cfg)
    with pytest.raises(L.GbpError, match="priors not set"):
        g.synchronous_iteration()
    with pytest.raises(ValueError):
        g.factors[0].eta_damping = 0.123
    g.close()


def _run_reference_script(script, args, timeout=600, env=None):
    """Run an UNMODIFIED script of the staged reference copy (baseline/_ref, put there by __graft_entry__.build()) against
    this engine: `python -m gbp_b200.run <script> ...` only puts gbp_b200/compat (packages gbp / utils / vis) first on sys.path."""
    import os
    import subprocess
    import sys
    from conftest import REF_COPY, ROOT
    path = os.path.join(REF_COPY, script)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: __graft_entry__.build() stages the reference copy (it ships to the GPU box with the tree)")
    import hashlib
    res = subprocess.run([sys.executable, "-m", "gbp_b200.run", path] + args, cwd=REF_COPY, env=dict(os.environ, PYTHONPATH=ROOT, **(env or {})),
                         capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout, hashlib.sha256(open(path, "rb").read()).hexdigest()


def _parse_ba_trace(out):
    import re
    rows = re.findall(r"Iteration (\d+) // ARE ([-+0-9.eE]+|nan|inf) // Energy ([-+0-9.eE]+|nan|inf) // Num factors relinearising (\d+)", out)
    return np.array([[float(x) for x in r] for r in rows])


# sha256 of /root/reference/ba.py at the survey commit (5670a49): the script the test runs must be the reference's own bytes
BA_PY_SHA256 = "ab5d4ac0748cadc8f562fa0e2e3be6a614664166effe66c6721f568e1fe7df86"


@pytest.mark.parametrize("data,fixture,extra", [
    ("fr1desk_vsmall.txt", "fr1desk_vsmall", []),
    ("fr1desk_vsmall.txt", "fr1desk_vsmall_huber", ["--loss", "huber"]),
    ("fr1desk_vsmall.txt", "fr1desk_vsmall_constant", ["--loss", "constant"]),
    ("fr1desk_vsmall.txt", "fr1desk_vsmall_float", ["--float_implementation"]),
    ("fr2robot2.txt", "fr2robot2", []),
    ("fr1desk.txt", "fr1desk", []),
])
def test_unmodified_reference_ba_py(data, fixture, extra):
    """BASELINE configs 2 / 3 as the north star states them: the reference's own ba.py, unmodified (`import vis`, the Python
    loops over graph.factors, viewer.update), runs on the GPU engine and prints the reference's trace to the printed digits
    (ba.py:68-105; ARE / energy to 4 decimals, relinearisation counts exactly on vsmall, within the branch-decision
    bound of the converged-belief test on fr1desk)."""
    G = load_golden(fixture)
    n_iters = int(G["n_iters"])
    out, sha = _run_reference_script("ba.py", ["--bal_file", f"data/{data}", "--n_iters", str(n_iters)] + extra)
    assert sha == BA_PY_SHA256
    C, Lm, F = len(G["in_cam0"]), len(G["in_lmk0"]), len(G["in_cam_id"])
    assert f"Number of keyframes: {C}" in out and f"Number of landmarks: {Lm}" in out and f"Number of measurement factors: {F}" in out
    tr = _parse_ba_trace(out)
    assert tr.shape == (n_iters, 4) and np.array_equal(tr[:, 0], np.arange(n_iters))
    float_impl = bool(G["float_impl"])
    # The --float_implementation run (100x weaker priors) is ill-conditioned: its last outer iteration (29) sits on an energy
    # spike (4.4e7 between neighbours of 1e5) that amplifies rounding differences by ~1e13 -- the g++ / libm build of the SAME
    # per-edge code is 1.2e-3 (full rows) / 1.6e-4 (factored rows) from the reference there and 3e-5 or better everywhere else
    # (tests/test_math_host.py), device builds landed between 2e-4 and 1.0e-3 depending on the rounding of sin / cos / rsqrt.
    # So: per iteration 5e-3 on that trace (1e-6 on the others), and the whole trace in the table norm at 1e-4 like the checkpoints.
    tol = 5e-3 if float_impl else 1e-6
    ref_are, ref_en = G["are"][:n_iters], G["energy"][:n_iters]
    assert np.all(np.abs(tr[:, 1] - ref_are) <= 0.5e-4 + tol * np.abs(ref_are))      # printed with 4 decimals
    assert np.all(np.abs(tr[:, 2] - ref_en) <= 0.5e-4 + tol * np.abs(ref_en))
    assert relerr(tr[:, 1], ref_are) < 1e-4 and relerr(tr[:, 2], ref_en) < 1e-4
    if float_impl:
        ok = np.abs(tr[:, 2] - ref_en) <= 0.5e-4 + 1e-4 * np.abs(ref_en)
        assert ok[:-1].all(), np.nonzero(~ok)[0]                                     # everything but the spike at 1e-4
    dn = np.abs(tr[:, 3].astype(int) - G["n_relin"][:n_iters])
    assert dn.max() <= (2 if fixture == "fr1desk" else 0), np.nonzero(dn)[0]
    if float_impl:
        assert out.count("Weakening priors") == 5


def test_unmodified_reference_ndim_posegraph():
    """BASELINE config 1: ndim_posegraph.py of the staged reference copy, unmodified, on the host graph classes (CPU by
    contract; GBP_LINEAR_DEVICE=0 keeps the graph off the GPU engine, which tests/test_lingraph.py covers) -- the check of
    tests/test_hostgraph.py, here from the shipped copy on the box."""
    G = load_golden("posegraph_n50_d3")
    out, _ = _run_reference_script("ndim_posegraph.py", ["--n_varnodes", "50", "--dim", "3"], env={"GBP_LINEAR_DEVICE": "0"})
    lines = [l for l in out.splitlines() if l.startswith("Iteration")]
    assert len(lines) == 50
    assert [float(l.split("Energy")[1].split("//")[0]) for l in lines] == G["energy"].tolist()
    assert [float(l.split("MAP")[1]) for l in lines] == G["dist"].tolist()


def test_snapshot_paths_agree():
    """Eager (fused graph + pinned snapshot) and lazy reads return the same numbers as explicit reads."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200 import _lib as L
    G = load_golden("fr1desk_vsmall")
    g = create_ba_graph(golden_problem(G), golden_configs(G))
    assert g._eager
    g.generate_priors_var(50.0); g.update_all_beliefs()
    for i in range(5):
        g.synchronous_iteration(robustify=True, local_relin=True)       # fused path, snapshot pending
        if i % 2:
            g.synchronous_iteration(robustify=True, local_relin=True)   # two in a row without reading
        a, e, n = g.metrics()
        sa, se, sn = g._eng.metrics()
        assert (a, e, n) == (sa / g._eng.F, se, sn)
        assert np.array_equal(g.cam_nodes[2].mu, g._eng.read(L.F_CAM_BELIEF)[2, 27:])
        assert np.array_equal(g.lmk_nodes[5].belief.eta, g._eng.read(L.F_LMK_BELIEF)[5, :3])
        g.lmk_nodes[5].mu = g.lmk_nodes[5].mu + 1e-3                    # client write while mirrors are hot
        g.factors[3].iters_since_relin = 4
    g._flush()                                                          # pending client writes reach the device
    assert g._eng.read(L.F_ITERS)[3, 0] == g.factors[3].iters_since_relin == 4
    g.reset()
    assert g._eng.read(L.F_MSG_CAM).max() == 0.0 and not g._snap_pending
    g.close()


def test_full_size_properties():
    """Size-independent properties on a large synthetic graph (2 M factors; the 10 M-factor configuration is the
    same code path and is exercised by bench.py): every belief is exactly prior + the sum of its incoming
    messages (linearity / checksum of checksums), messages are symmetric rank-2 PSD, means solve Lambda mu = eta,
    every factor relinearises on schedule, and the result does not depend on the engine's storage order."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.engine import unpack_sym
    from gbp_b200.synthetic import make_synthetic
    from gbp_b200 import _lib as L
    prob = make_synthetic(200, 200_000, 10, seed=1)
    cfg = dict(gauss_noise_std=2, loss="huber", Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
               eta_damping=0.4)
    g = create_ba_graph(prob, cfg)
    assert g._eng.F == 2_000_000 and g._eng.tile_edges == 64
    # large graphs get the bandwidth-oriented build: factored keyframe messages, early issue, L2 prefetch 38 k edges ahead
    assert (g._eng.msg_cam_width, g._eng.sweep_variant, g._eng.prefetch_tiles) == (18, 2, 600)
    g.generate_priors_var(50.0)
    g.update_all_beliefs()
    g.iterate(12, robustify=True, local_relin=True)
    e = g._eng
    adj = e.read(L.F_ADJ)
    mc, ml = e.read(L.F_MSG_CAM), e.read(L.F_MSG_LMK)
    cb, lb = e.read(L.F_CAM_BELIEF), e.read(L.F_LMK_BELIEF)
    cp, lp = e.read(L.F_CAM_PRIOR), e.read(L.F_LMK_PRIOR)
    # (1) beliefs = prior + sum of incoming messages, for ALL variables (float64 sums in a different order)
    cs = cp.copy(); np.add.at(cs, adj[:, 0], mc)
    ls = lp.copy(); np.add.at(ls, adj[:, 1], ml)
    assert relerr(cb[:, :27], cs) < 1e-12 and relerr(lb[:, :9], ls) < 1e-12
    # (2) means solve the belief system
    lam_c, lam_l = unpack_sym(cb[:, 6:27], 6), unpack_sym(lb[:, 3:9], 3)
    assert relerr(np.einsum("vij,vj->vi", lam_c, cb[:, 27:]), cb[:, :6]) < 1e-9
    assert relerr(np.einsum("vij,vj->vi", lam_l, lb[:, 9:]), lb[:, :3]) < 1e-9
    # (3) messages: symmetric PSD of rank <= 2 (sample)
    idx = np.linspace(0, e.F - 1, 5000).astype(int)
    ev = np.linalg.eigvalsh(unpack_sym(mc[idx, 6:], 6))
    assert ev.min() > -1e-9 * ev.max() and np.all(ev[:, :4].max(axis=1) < 1e-9 * ev[:, 5] + 1e-300)
    ev3 = np.linalg.eigvalsh(unpack_sym(ml[idx, 3:], 3))
    assert ev3.min() > -1e-9 * ev3.max() and np.all(ev3[:, 0] < 1e-9 * ev3[:, 2] + 1e-300)
    # (4) relinearisation schedule without client resets: iters_since_relin starts at 1, reaches min_linear_iters = 8
    #     before sweep 7, the factor relinearises there (-> 0) and counts 4 more sweeps; factors that never moved read 13
    its = e.read(L.F_ITERS)[:, 0]
    #     (a factor whose mean had moved less than beta at sweep 7 relinearises one to three sweeps later)
    assert set(np.unique(its).tolist()) <= {0, 1, 2, 3, 4, 13} and (its == 4).mean() > 0.9
    assert np.isfinite(cb).all() and np.isfinite(lb).all() and np.isfinite(mc).all()
    # (5) storage order (tiles, landmark blocks) and kernel variant do not change the state
    h = create_ba_graph(prob, cfg, tile_edges=128, lmk_block=50_000, kernel_variant=1)
    h.generate_priors_var(50.0)
    h.update_all_beliefs()
    h.iterate(12, robustify=True, local_relin=True)
    assert relerr(h._eng.read(L.F_LMK_BELIEF), lb) < 1e-9 and relerr(h._eng.read(L.F_CAM_BELIEF), cb) < 1e-9
    a1, a2 = g.are(), h.are()
    assert abs(a1 - a2) < 1e-9 * a1 and a1 < 5.0
    g.close(); h.close()


def test_synthetic_small_against_reference_fixture():
    """BASELINE configs 4-5 are pinned on a down-scaled instance of the same generator: the unmodified reference
    ran `make_synthetic(20, 2000, 10, seed=0)` for 30 iterations (tests/golden/synth_small.npz)."""
    from gbp_b200.ba import create_ba_graph
    from gbp_b200.synthetic import make_synthetic
    G = load_golden("synth_small")
    prob = make_synthetic(20, 2000, 10, seed=0)
    ref = golden_problem(G)
    for k in ("cam_id", "lmk_id", "z", "cam_means", "lmk_means", "K4"):      # the generator is deterministic
        assert np.array_equal(getattr(prob, k), getattr(ref, k)), k
    graph = create_ba_graph(prob, golden_configs(G), tile_edges=64, lmk_block=700)
    cks = set(G["checkpoints"].tolist())

    def on_iter(i):
        if i in cks:
            _check_snapshot(graph, G, f"s{i}", TOL_EARLY if i <= 1 else 1e-5)

    are, en, nrel = _run_loop(graph, G, int(G["n_iters"]), on_iter)
    assert np.array_equal(nrel, G["n_relin"])
    assert relerr(are, G["are"]) < 1e-6 and relerr(en, G["energy"]) < 1e-6
    graph.close()
