"""BASELINE config 1 (CPU plumbing): the generic host FactorGraph with Python-callable factors.
The problem construction below restates ndim_posegraph.py:35-90 (same NumPy legacy random stream) so the
test also runs where the reference checkout is absent; the expected trace is the reference's own output."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

COMPAT = os.path.join(ROOT, "gbp_b200", "compat")


def build_posegraph(n_varnodes, dim, M, noise_std):
    sys.path[:0] = [p for p in (COMPAT, ROOT) if p not in sys.path]
    from gbp_b200 import hostgraph as gbp
    from gbp_b200.compat.gbp.factors import linear_displacement  # noqa: F401
    np.random.seed(0)
    priors_mu = np.random.rand(n_varnodes, dim) * 10
    prior_lambda = np.linalg.inv(3 * np.eye(dim))
    priors_eta = [prior_lambda @ mu for mu in priors_mu]
    meas, ids = [], []
    for i, mu in enumerate(priors_mu):
        dists = np.array([np.linalg.norm(mu - m1) for m1 in priors_mu])
        for j in dists.argsort()[1:M + 1]:
            if [j, i] not in ids:
                meas.append(mu - priors_mu[j] + np.random.normal(0., noise_std, dim))
                ids.append([i, j])
    graph = gbp.FactorGraph(nonlinear_factors=False)
    for i in range(n_varnodes):
        v = gbp.VariableNode(i, dim)
        v.prior.eta, v.prior.lam = priors_eta[i], prior_lambda
        graph.var_nodes.append(v)
    for f, z in enumerate(meas):
        a, b = graph.var_nodes[ids[f][0]], graph.var_nodes[ids[f][1]]
        fac = gbp.Factor(f, [a, b], z, noise_std, linear_displacement.meas_fn, linear_displacement.jac_fn,
                         loss=None, mahalanobis_threshold=2)
        a.adj_factors.append(fac); b.adj_factors.append(fac)
        graph.factors.append(fac)
    graph.update_all_beliefs()
    graph.compute_all_factors()
    return graph


def test_config1_trace_matches_reference_output():
    G = load_golden("posegraph_n50_d3")
    graph = build_posegraph(50, 3, 10, 1.0)
    mu, _ = graph.joint_distribution_cov()
    for i in range(50):
        graph.synchronous_iteration()
        assert f"{graph.energy():.4f}" == f"{G['energy'][i]:.4f}", i
        assert f"{np.linalg.norm(graph.get_means() - mu):4f}" == f"{G['dist'][i]:4f}", i


def test_gbp_converges_to_the_batch_solution_on_a_tree():
    """On a tree (M = 1 chain-like graph) GBP means converge to the dense MAP solution (gbp/gbp.py:136-144)."""
    graph = build_posegraph(30, 3, 2, 1.0)
    mu, sigma = graph.joint_distribution_cov()
    for _ in range(60):
        graph.synchronous_iteration()
    assert np.linalg.norm(graph.get_means() - mu) < 1e-3
    eta, lam = graph.joint_distribution_inf()
    assert np.allclose(lam, lam.T) and np.allclose(np.linalg.solve(lam, eta), mu)


def test_nonlinear_host_factor_with_robust_loss_and_relinearisation():
    """The generic graph also carries the nonlinear machinery (robustify / relinearise / damping)."""
    from gbp_b200 import hostgraph as gbp
    from gbp_b200.se3 import reprojection_jac_fn, reprojection_meas_fn, check_jac
    K = np.array([[517.3, 0., 318.6], [0., 516.5, 255.3], [0., 0., 1.]])
    g = gbp.FactorGraph(nonlinear_factors=True, eta_damping=0.4, beta=0.01, num_undamped_iters=2, min_linear_iters=3)
    cam, lmk = gbp.VariableNode(0, 6), gbp.VariableNode(1, 3)
    cam.mu, lmk.mu = np.array([0.05, -0.02, 0.1, 0.1, -0.05, 0.02]), np.array([0.2, -0.1, 2.0])
    cam.prior.lam, lmk.prior.lam = np.eye(6) * 1e4, np.eye(3) * 1.0
    cam.prior.eta, lmk.prior.eta = cam.prior.lam @ cam.mu, lmk.prior.lam @ lmk.mu
    g.var_nodes = [cam, lmk]
    x_true = np.concatenate([cam.mu, [0.25, -0.05, 2.1]])
    f = gbp.Factor(0, [cam, lmk], reprojection_meas_fn(x_true, K), 1.0, reprojection_meas_fn, reprojection_jac_fn, "huber", 2.0, K)
    cam.adj_factors.append(f); lmk.adj_factors.append(f); g.factors.append(f)
    f.compute_factor(np.concatenate([cam.mu, lmk.mu]))
    g.update_all_beliefs()
    e0 = g.energy()
    for _ in range(12):
        g.synchronous_iteration(robustify=True, local_relin=True)
    assert g.energy() < 1e-2 * e0 and f.iters_since_relin < 12
    check_jac(reprojection_jac_fn, x_true, reprojection_meas_fn, K)


def test_unmodified_reference_script_runs_if_present():
    script = "/root/reference/ndim_posegraph.py"
    if not os.path.exists(script):
        pytest.skip("reference checkout not present")
    out = subprocess.run([sys.executable, "-m", "gbp_b200.run", script, "--n_varnodes", "50", "--dim", "3"], cwd=ROOT,
                         capture_output=True, text=True, check=True).stdout
    G = load_golden("posegraph_n50_d3")
    lines = [l for l in out.splitlines() if l.startswith("Iteration")]
    assert len(lines) == 50
    assert [float(l.split("Energy")[1].split("//")[0]) for l in lines] == G["energy"].tolist()
    assert [float(l.split("MAP")[1]) for l in lines] == G["dist"].tolist()
