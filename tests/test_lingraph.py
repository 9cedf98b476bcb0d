"""Generic linear pairwise graphs (the path of ndim_posegraph.py): host logic of the device backend on CPU,
and under -m gpu the device engine (gbp_lin_* through ctypes) against the host NumPy classes -- which reproduce the unmodified
reference script to the printed digit (tests/test_hostgraph.py) -- and against the reference's own traces (tests/golden/posegraph_*)."""
import numpy as np
import pytest

from conftest import load_golden, relerr


def _displacement(x):          # gbp/factors/linear_displacement.py:8-14
    d = len(x) // 2
    return np.hstack((-np.eye(d), np.eye(d)))


def _build(n_vars=20, dim=3, M=4, seed=0, damping=0.0, meas_rows=None):
    """The construction of ndim_posegraph.py:34-90 (nearest-neighbour displacement factors), seeded locally."""
    from gbp_b200 import hostgraph as hg
    rng = np.random.default_rng(seed)
    mus = rng.random((n_vars, dim)) * 10
    g = hg.FactorGraph(nonlinear_factors=False, eta_damping=damping)
    for i in range(n_vars):
        v = hg.VariableNode(i, dim)
        v.prior.lam = np.linalg.inv(3 * np.eye(dim))
        v.prior.eta = v.prior.lam @ mus[i]
        g.var_nodes.append(v)
    pairs = []
    for i in range(n_vars):
        for j in np.argsort(np.linalg.norm(mus - mus[i], axis=1))[1:M + 1]:
            if [j, i] not in pairs and [i, j] not in pairs:
                pairs.append([i, int(j)])
    rows = dim if meas_rows is None else meas_rows
    A = rng.standard_normal((rows, 2 * dim)) if meas_rows is not None else None
    for f, (i, j) in enumerate(pairs):
        if A is None:
            jac = _displacement
            meas = lambda x: _displacement(x) @ x          # noqa: E731
            z = mus[i] - mus[j] + rng.normal(0, 1.0, dim)
        else:
            jac = lambda x, A=A: A                          # noqa: E731
            meas = lambda x, A=A: A @ x + 0.5               # noqa: E731  (affine: h(0) != 0)
            z = A @ np.concatenate([mus[i], mus[j]]) + 0.5 + rng.normal(0, 1.0, rows)
        fac = hg.Factor(f, [g.var_nodes[i], g.var_nodes[j]], z, 1.0 + 0.1 * (f % 3), meas, jac, loss=None, mahalanobis_threshold=2)
        g.var_nodes[i].adj_factors.append(fac); g.var_nodes[j].adj_factors.append(fac); g.factors.append(fac)
    return g


def test_tables_from_host_graph_cpu():
    from gbp_b200 import hostgraph as hg, lingraph
    hg.USE_DEVICE = False
    try:
        g = _build(12, 3, 3, seed=1)
        g.update_all_beliefs(); g.compute_all_factors()
        t = lingraph.tables_from_host_graph(g)
        F = len(g.factors)
        assert t["dim"] == 3 and t["J"].shape == (F, 3, 6) and t["adj_ptr"][-1] == 2 * F and sorted(t["adj_msg"].tolist()) == list(range(2 * F))
        for k, f in enumerate(g.factors):
            np.testing.assert_allclose(t["J"][k], _displacement(np.zeros(6)))
            np.testing.assert_allclose(t["b"][k], f.measurement)                 # linear: J x0 + z - h(x0) = z
            np.testing.assert_allclose(f.factor.lam, t["J"][k].T @ t["J"][k] / t["var"][k])
            np.testing.assert_allclose(f.factor.eta, t["J"][k].T @ t["b"][k] / t["var"][k])
        for k, v in enumerate(g.var_nodes):                                      # adj_factors order is the summation order
            got = t["adj_msg"][t["adj_ptr"][k]:t["adj_ptr"][k + 1]]
            want = [2 * g.factors.index(f) + [id(a) for a in f.adj_var_nodes].index(id(v)) for f in v.adj_factors]
            assert got.tolist() == want
        # not eligible: nonlinear graph, unequal dofs, robust loss, a factor missing from an adjacency list
        g.nonlinear_factors = True
        assert lingraph.tables_from_host_graph(g) is None
        g.nonlinear_factors = False
        g.factors[0].loss = "huber"
        assert lingraph.tables_from_host_graph(g) is None
        g.factors[0].loss = None
        g.var_nodes[0].adj_factors.pop()
        assert lingraph.tables_from_host_graph(g) is None
        g2 = _build(6, 2, 2)
        g2.var_nodes[1].dofs = 3
        assert lingraph.tables_from_host_graph(g2) is None
    finally:
        hg.USE_DEVICE = None


def test_host_fallback_without_a_device_cpu():
    """Automatic mode stays on the host when no CUDA device is visible; USE_DEVICE = True refuses loudly."""
    from gbp_b200 import hostgraph as hg, lingraph
    if lingraph.device_available():
        pytest.skip("a CUDA device is present")
    g = _build(8, 2, 2)
    g.update_all_beliefs(); g.compute_all_factors()
    g.synchronous_iteration()
    assert g._dev is None and np.isfinite(g.energy())
    hg.USE_DEVICE = True
    try:
        h = _build(8, 2, 2)
        h.update_all_beliefs(); h.compute_all_factors()
        with pytest.raises(RuntimeError):
            h.synchronous_iteration()
    finally:
        hg.USE_DEVICE = None


@pytest.mark.gpu
@pytest.mark.parametrize("dim,damping,meas_rows", [(1, 0.0, None), (2, 0.0, None), (3, 0.3, None), (6, 0.0, None), (4, 0.2, 2), (3, 0.0, 3)])
def test_device_graph_equals_host_graph(dim, damping, meas_rows):
    """Same graph on the host classes and on the device engine: beliefs, messages, energy and means through 25 iterations;
    joint_distribution_inf / _cov from the device (dense Cholesky) against NumPy."""
    from gbp_b200 import hostgraph as hg
    hg.USE_DEVICE = False
    a = _build(30, dim, 4, seed=dim, damping=damping, meas_rows=meas_rows)
    a.update_all_beliefs(); a.compute_all_factors()
    mu_a, sig_a = a.joint_distribution_cov()
    eta_a, lam_a = a.joint_distribution_inf()
    hg.USE_DEVICE = True
    try:
        b = _build(30, dim, 4, seed=dim, damping=damping, meas_rows=meas_rows)
        b.update_all_beliefs(); b.compute_all_factors()
        mu_b, sig_b = b.joint_distribution_cov()
        eta_b, lam_b = b.joint_distribution_inf()
        assert b._dev is not None
        assert relerr(eta_b, eta_a) < 1e-12 and relerr(lam_b, lam_a) < 1e-12
        assert relerr(mu_b, mu_a) < 1e-9 and relerr(sig_b, sig_a) < 1e-9
        for it in range(25):
            hg.USE_DEVICE = False
            a.synchronous_iteration()
            hg.USE_DEVICE = True
            b.synchronous_iteration()
            assert abs(b.energy() - a.energy()) < 1e-9 * max(a.energy(), 1.0), it
        assert relerr(b.get_means(), a.get_means()) < 1e-9
        for va, vb in zip(a.var_nodes, b.var_nodes):           # the node / factor objects follow the device state
            assert relerr(vb.belief.lam, va.belief.lam) < 1e-9 and relerr(vb.belief.eta, va.belief.eta) < 1e-9 and relerr(vb.mu, va.mu) < 1e-9
        for fa, fb in zip(a.factors, b.factors):
            for s in (0, 1):
                assert relerr(fb.messages[s].lam, fa.messages[s].lam) < 1e-9 and relerr(fb.messages[s].eta, fa.messages[s].eta) < 1e-9
        assert b._dev.launch_count() >= 50
        # a host-only call takes the state back from the device and continues identically
        b.compute_all_messages(); a.compute_all_messages()
        assert b._dev is None
        b.update_all_beliefs(); a.update_all_beliefs()
        assert relerr(b.get_means(), a.get_means()) < 1e-9
    finally:
        hg.USE_DEVICE = None


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,args", [("posegraph_n50_d3", ["--n_varnodes", "50", "--dim", "3"]), ("posegraph_default", [])])
def test_unmodified_ndim_posegraph_on_the_device(fixture, args):
    """BASELINE config 1's script, unmodified, with its FactorGraph on the GPU engine (automatic on a box with a device): energy and
    distance-to-MAP traces of the reference to the printed precision (the MAP itself is the device's dense Cholesky solve)."""
    from test_ba_gpu import _run_reference_script
    G = load_golden(fixture)
    out, _ = _run_reference_script("ndim_posegraph.py", args, env={"GBP_LINEAR_DEVICE": "1"})
    lines = [l for l in out.splitlines() if l.startswith("Iteration")]
    assert len(lines) == len(G["energy"])
    energy = np.array([float(l.split("Energy")[1].split("//")[0]) for l in lines])
    dist = np.array([float(l.split("MAP")[1]) for l in lines])
    assert np.all(np.abs(energy - G["energy"]) <= 1.01e-4) and np.all(np.abs(dist - G["dist"]) <= 1.01e-6)
