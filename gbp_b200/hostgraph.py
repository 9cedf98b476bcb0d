"""Generic factor graph with arbitrary Python measurement callables.

Mirror of the reference's ``gbp/gbp.py`` classes ``FactorGraph`` / ``VariableNode`` / ``Factor``
for graphs the client assembles object by object (BASELINE config 1: ``ndim_posegraph.py``,
linear displacement factors built from Python callables).  The arithmetic here is host NumPy --
config 1 is CPU plumbing by contract -- but a graph of pairwise LINEAR factors between variables
of equal dimension <= 6 moves onto the GPU engine of ``gbp_b200.lingraph``
when a CUDA device is present: ``USE_DEVICE`` = None (automatic) / True (required) / False (host).
The bundle-adjustment path never uses this module: reprojection graphs are built by
``gbp_b200.ba`` and run on the GPU.

Same constructor signatures, attribute names and update rules as the reference:
  synchronous_iteration   gbp/gbp.py:86-92      robustify_loss    gbp/gbp.py:296-332
  relinearise_factors     gbp/gbp.py:64-80      compute_factor    gbp/gbp.py:267-294
  compute_messages        gbp/gbp.py:334-373    update_belief     gbp/gbp.py:176-198
  joint_distribution_*    gbp/gbp.py:94-144     energy            gbp/gbp.py:36-44
"""
from __future__ import annotations

import numpy as np

from .gaussian import NdimGaussian

import os as _os

# None: device when available and the graph qualifies; True: must; False: never.  An UNMODIFIED client script cannot set this, so
# the launcher honours GBP_LINEAR_DEVICE=0 / 1 (python -m gbp_b200.run ndim_posegraph.py).
USE_DEVICE = {"0": False, "1": True}.get(_os.environ.get("GBP_LINEAR_DEVICE", ""), None)
SYNC_OBJECTS_MAX = 4096    # graphs up to this many variables / factors refresh the node / factor objects after every device call


def _mean_of(g):
    return np.linalg.inv(g.lam) @ g.eta


class VariableNode:
    def __init__(self, variable_id, dofs):
        self.variableID = variable_id
        self.dofs = dofs
        self.adj_factors = []
        self.mu = np.zeros(dofs)
        self.Sigma = np.zeros([dofs, dofs])
        self.belief = NdimGaussian(dofs)
        self.prior = NdimGaussian(dofs)
        self.prior_lambda_end = -1
        self.prior_lambda_logdiff = -1

    def update_belief(self):
        """Product of the prior and every incoming message; then broadcast to adjacent factors."""
        eta, lam = np.array(self.prior.eta, dtype=float), np.array(self.prior.lam, dtype=float)
        slots = [(f, f.adj_vIDs.index(self.variableID)) for f in self.adj_factors]
        for f, k in slots:
            eta += f.messages[k].eta
            lam += f.messages[k].lam
        self.belief.eta, self.belief.lam = eta, lam
        self.Sigma = np.linalg.inv(lam)
        self.mu = self.Sigma @ eta
        for f, k in slots:
            f.adj_beliefs[k].eta, f.adj_beliefs[k].lam = eta, lam


class Factor:
    def __init__(self, factor_id, adj_var_nodes, measurement, gauss_noise_std, meas_fn, jac_fn, loss=None,
                 mahalanobis_threshold=2, *args):
        self.factorID = factor_id
        self.adj_var_nodes = adj_var_nodes
        self.adj_vIDs = [v.variableID for v in adj_var_nodes]
        self.adj_beliefs = [NdimGaussian(v.dofs) for v in adj_var_nodes]
        self.messages = [NdimGaussian(v.dofs) for v in adj_var_nodes]
        self.dofs_conditional_vars = sum(v.dofs for v in adj_var_nodes)
        self.factor = NdimGaussian(self.dofs_conditional_vars)
        self.linpoint = np.zeros(self.dofs_conditional_vars)
        self.measurement = measurement
        self.gauss_noise_var = gauss_noise_std ** 2
        self.meas_fn, self.jac_fn, self.args = meas_fn, jac_fn, args
        self.adaptive_gauss_noise_var = gauss_noise_std ** 2
        self.loss = loss
        self.mahalanobis_threshold = mahalanobis_threshold
        self.robust_flag = False
        self.eta_damping = 0.
        self.iters_since_relin = 1

    # -- helpers
    def _adj_means(self):
        parts = [_mean_of(b) for b in self.adj_beliefs]
        return np.concatenate(parts) if parts else np.array([])

    def _offsets(self):
        return np.concatenate([[0], np.cumsum([v.dofs for v in self.adj_var_nodes])]).astype(int)

    def compute_residual(self):
        return self.meas_fn(self._adj_means(), *self.args) - self.measurement

    def energy(self):
        return 0.5 * np.linalg.norm(self.compute_residual()) ** 2 / self.adaptive_gauss_noise_var

    def compute_factor(self, linpoint=None, update_self=True):
        self.linpoint = list(self._adj_means()) if linpoint is None else linpoint
        J = self.jac_fn(self.linpoint, *self.args)
        pred = self.meas_fn(self.linpoint, *self.args)
        w = 1.0 / self.adaptive_gauss_noise_var
        if isinstance(self.measurement, float):
            lam = w * np.outer(J, J)
            eta = w * J.T * (J @ self.linpoint + self.measurement - pred)
        else:
            W = np.eye(len(self.measurement)) * w
            lam = J.T @ W @ J
            eta = (J.T @ W) @ (J @ self.linpoint + self.measurement - pred)
        if update_self:
            self.factor.eta, self.factor.lam = eta, lam
        return eta, lam

    def robustify_loss(self):
        old = self.adaptive_gauss_noise_var
        if self.loss is None:
            self.adaptive_gauss_noise_var = self.gauss_noise_var
        elif self.loss in ("huber", "constant"):
            pred = self.meas_fn(self.linpoint, *self.args)
            m = np.linalg.norm(self.measurement - pred) / np.sqrt(self.gauss_noise_var)
            n = self.mahalanobis_threshold
            self.robust_flag = bool(m > n)
            if not self.robust_flag:
                self.adaptive_gauss_noise_var = self.gauss_noise_var
            elif self.loss == "huber":
                self.adaptive_gauss_noise_var = self.gauss_noise_var * m ** 2 / (2 * (n * m - 0.5 * n ** 2))
            else:
                self.adaptive_gauss_noise_var = m ** 2
        scale = old / self.adaptive_gauss_noise_var
        self.factor.eta *= scale
        self.factor.lam *= scale

    def compute_messages(self, eta_damping):
        off = self._offsets()
        n = len(self.adj_vIDs)
        new = []
        for v in range(n):
            eta, lam = self.factor.eta.copy(), self.factor.lam.copy()
            for u in range(n):          # variable->factor messages of the other variables
                if u != v:
                    s = slice(off[u], off[u + 1])
                    eta[s] += self.adj_beliefs[u].eta - self.messages[u].eta
                    lam[s, s] += self.adj_beliefs[u].lam - self.messages[u].lam
            keep = np.arange(off[v], off[v + 1])
            drop = np.concatenate([np.arange(0, off[v]), np.arange(off[v + 1], off[-1])])
            lam_kd = lam[np.ix_(keep, drop)]
            inv_dd = np.linalg.inv(lam[np.ix_(drop, drop)])
            m_lam = lam[np.ix_(keep, keep)] - lam_kd @ inv_dd @ lam[np.ix_(drop, keep)]
            m_eta = eta[keep] - lam_kd @ inv_dd @ eta[drop]
            new.append(((1 - eta_damping) * m_eta + eta_damping * self.messages[v].eta, m_lam))
        for v, (e, l) in enumerate(new):
            self.messages[v].eta, self.messages[v].lam = e, l


class FactorGraph:
    def __init__(self, nonlinear_factors=True, eta_damping=0.0, beta=None, num_undamped_iters=None,
                 min_linear_iters=None):
        self.var_nodes, self.factors = [], []
        self.n_var_nodes = self.n_factor_nodes = self.n_edges = 0
        self.nonlinear_factors = nonlinear_factors
        self.eta_damping = eta_damping
        if nonlinear_factors:
            self.beta = beta
            self.num_undamped_iters = num_undamped_iters
            self.min_linear_iters = min_linear_iters
        self._dev = None            # gbp_b200.lingraph.LinearGraphEngine once the graph runs on the GPU
        self._dev_tried = False

    # ------------------------------------------------------------------ device backend (linear pairwise graphs)
    def device_backend(self):
        """The GPU engine of this graph, built from the objects on first use (after compute_all_factors: the factors'
        linearisation points must exist), or None when the graph stays on the host."""
        if self._dev is not None or self._dev_tried or USE_DEVICE is False:
            return self._dev
        self._dev_tried = True
        from . import lingraph
        if not lingraph.device_available():
            if USE_DEVICE:
                raise RuntimeError("hostgraph.USE_DEVICE is True but no CUDA device / library is available")
            return None
        t = lingraph.tables_from_host_graph(self)
        if t is None:
            if USE_DEVICE:
                raise RuntimeError("hostgraph.USE_DEVICE is True but this graph is not a pairwise linear graph of equal dofs <= 6")
            return None
        msg_eta, msg_lam = t.pop("msg_eta"), t.pop("msg_lam")
        self._dev = lingraph.LinearGraphEngine(eta_damping=self.eta_damping, **t)
        if np.any(msg_eta) or np.any(msg_lam):
            self._dev.set_messages(msg_eta, msg_lam)
        self._dev.update_beliefs()
        return self._dev

    def _objects_from_device(self, force=False):
        """Beliefs, means and messages of the device graph back into the VariableNode / Factor objects."""
        d = self._dev
        if d is None or (not force and max(len(self.var_nodes), len(self.factors)) > SYNC_OBJECTS_MAX):
            return
        eta, lam, mu = d.beliefs()
        for k, v in enumerate(self.var_nodes):
            v.belief.eta, v.belief.lam, v.mu = eta[k].copy(), lam[k].copy(), mu[k].copy()
            v.Sigma = np.linalg.inv(lam[k])
        meta, mlam = d.messages()
        index = {id(v): k for k, v in enumerate(self.var_nodes)}
        for k, f in enumerate(self.factors):
            for side in (0, 1):
                f.messages[side].eta, f.messages[side].lam = meta[2 * k + side].copy(), mlam[2 * k + side].copy()
                vk = index[id(f.adj_var_nodes[side])]
                f.adj_beliefs[side].eta, f.adj_beliefs[side].lam = eta[vk], lam[vk]

    def _leave_device(self):
        """Back to the host arithmetic (a host-only method was called): the objects take over the device state."""
        if self._dev is not None:
            self._objects_from_device(force=True)
            self._dev.close()
            self._dev = None
        self._dev_tried = False

    def energy(self):
        if self._dev is not None:
            return self._dev.energy()
        return sum(f.energy() for f in self.factors)

    def compute_all_messages(self, local_relin=True):
        self._leave_device()
        local = self.nonlinear_factors and local_relin
        for f in self.factors:
            if local:
                if f.iters_since_relin == self.num_undamped_iters:
                    f.eta_damping = self.eta_damping
                f.compute_messages(f.eta_damping)
            else:
                f.compute_messages(self.eta_damping)

    def update_all_beliefs(self):
        if self._dev is not None:
            self._dev.update_beliefs()
            self._objects_from_device()
            return
        for v in self.var_nodes:
            v.update_belief()

    def compute_all_factors(self):
        self._leave_device()          # new linearisation points / factors: the device tables are rebuilt on the next sweep
        for f in self.factors:
            f.compute_factor()

    def relinearise_factors(self):
        if not self.nonlinear_factors:
            return
        self._leave_device()
        for f in self.factors:
            means = f._adj_means()
            if np.linalg.norm(f.linpoint - means) > self.beta and f.iters_since_relin >= self.min_linear_iters:
                f.compute_factor(linpoint=means)
                f.iters_since_relin = 0
                f.eta_damping = 0.0
            else:
                f.iters_since_relin += 1

    def robustify_all_factors(self):
        self._leave_device()
        for f in self.factors:
            f.robustify_loss()

    def synchronous_iteration(self, local_relin=True, robustify=False):
        if not robustify and not self.nonlinear_factors and self.device_backend() is not None:
            self._dev.iterate(1)              # gbp/gbp.py:86-92 for a linear graph: messages (graph-level damping), beliefs
            self._objects_from_device()
            return
        if robustify:
            self.robustify_all_factors()
        if self.nonlinear_factors and local_relin:
            self.relinearise_factors()
        self.compute_all_messages(local_relin=local_relin)
        self.update_all_beliefs()

    def joint_distribution_inf(self):
        """Joint information form over all variables (priors + factors at their linearisation points)."""
        if not self.nonlinear_factors and self.device_backend() is not None:
            _, _, eta, lam = self._dev.joint_solve(want_sigma=False, want_inf=True)
            return eta, lam
        dofs = {v.variableID: v.dofs for v in self.var_nodes}
        start, tot = {}, 0
        for v in self.var_nodes:
            start[v.variableID] = tot
            tot += v.dofs
        eta, lam = np.zeros(tot), np.zeros((tot, tot))
        for v in self.var_nodes:
            s = slice(start[v.variableID], start[v.variableID] + v.dofs)
            eta[s] += v.prior.eta
            lam[s, s] += v.prior.lam
        for f in self.factors:
            ix = np.concatenate([np.arange(start[i], start[i] + dofs[i]) for i in f.adj_vIDs])
            eta[ix] += f.factor.eta
            lam[np.ix_(ix, ix)] += f.factor.lam
        return eta, lam

    def joint_distribution_cov(self):
        if not self.nonlinear_factors and self.device_backend() is not None:
            mu, sigma, _, _ = self._dev.joint_solve(want_sigma=True)       # dense Cholesky solve + inverse on the device
            return mu, sigma
        eta, lam = self.joint_distribution_inf()
        sigma = np.linalg.inv(lam)
        return sigma @ eta, sigma

    def get_means(self):
        if self._dev is not None:
            return self._dev.means().ravel()
        return np.concatenate([v.mu for v in self.var_nodes]) if self.var_nodes else np.array([])
