"""Device engine for graphs of pairwise LINEAR factors between variables of equal dimension (ctypes over gbp_lin_*).

This is the GPU path of the generic ``gbp.FactorGraph(nonlinear_factors=False)`` that ``ndim_posegraph.py`` builds
(SURVEY section 8(f) rank 4).  `LinearGraphEngine` is the raw engine; `hostgraph.FactorGraph` moves a graph the client
built object by object onto it (`device_backend`) and keeps the reference's object API in step.

Reference: FactorGraph.synchronous_iteration gbp/gbp.py:86-92 (compute_all_messages :46-54, 334-373; update_all_beliefs
:56-58, 176-198), energy :36-44, joint_distribution_inf / _cov :94-144, Factor.compute_factor :267-294.
There is no CPU implementation behind this class: without the CUDA library / a device it raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

MAX_DOFS = 6
BELIEFS, MESSAGES = 0, 1


class LinearGraphEngine:
    def __init__(self, dim, var_i, var_j, J, b, var, prior_eta, prior_lam, adj_ptr, adj_msg, eta_damping=0.0, device=0, stream=None):
        lib = L.load()
        self._lib, self._h = lib, None
        self.dim = int(dim)
        self.var_i = np.ascontiguousarray(var_i, dtype=np.int32)
        self.var_j = np.ascontiguousarray(var_j, dtype=np.int32)
        self.F = len(self.var_i)
        prior_eta = np.ascontiguousarray(prior_eta, dtype=np.float64).reshape(-1, self.dim)
        self.V = len(prior_eta)
        prior_lam = np.ascontiguousarray(prior_lam, dtype=np.float64).reshape(self.V, self.dim, self.dim)
        J = np.ascontiguousarray(J, dtype=np.float64).reshape(self.F, self.dim, 2 * self.dim)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(self.F, self.dim)
        var = np.ascontiguousarray(var, dtype=np.float64).reshape(self.F)
        adj_ptr = np.ascontiguousarray(adj_ptr, dtype=np.int32)
        adj_msg = np.ascontiguousarray(adj_msg, dtype=np.int32)
        h = C.c_void_p()
        L.check(lib.gbp_lin_create(self.dim, self.V, self.F, L.ptr(self.var_i), L.ptr(self.var_j), L.ptr(J), L.ptr(b), L.ptr(var),
                                   L.ptr(prior_eta), L.ptr(prior_lam), L.ptr(adj_ptr), L.ptr(adj_msg), float(eta_damping), int(device),
                                   C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            self._lib.gbp_lin_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_messages(self, eta, lam):
        eta = np.ascontiguousarray(eta, dtype=np.float64).reshape(2 * self.F, self.dim)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(2 * self.F, self.dim, self.dim)
        L.check(self._lib.gbp_lin_set_messages(self._h, L.ptr(eta), L.ptr(lam)))

    def update_beliefs(self):
        L.check(self._lib.gbp_lin_update_beliefs(self._h))

    def iterate(self, n_iters=1):
        L.check(self._lib.gbp_lin_iterate(self._h, int(n_iters)))

    def energy(self):
        out = C.c_double()
        L.check(self._lib.gbp_lin_energy(self._h, C.byref(out)))
        return out.value

    def beliefs(self):
        eta = np.empty((self.V, self.dim)); lam = np.empty((self.V, self.dim, self.dim)); mu = np.empty((self.V, self.dim))
        L.check(self._lib.gbp_lin_read(self._h, BELIEFS, L.ptr(eta), L.ptr(lam), L.ptr(mu)))
        return eta, lam, mu

    def means(self):
        mu = np.empty((self.V, self.dim))
        L.check(self._lib.gbp_lin_read(self._h, BELIEFS, None, None, L.ptr(mu)))
        return mu

    def messages(self):
        eta = np.empty((2 * self.F, self.dim)); lam = np.empty((2 * self.F, self.dim, self.dim))
        L.check(self._lib.gbp_lin_read(self._h, MESSAGES, L.ptr(eta), L.ptr(lam), None))
        return eta, lam

    def joint_solve(self, want_sigma=True, want_inf=False):
        n = self.V * self.dim
        mu = np.empty(n)
        sigma = np.empty((n, n)) if want_sigma else None
        eta = np.empty(n) if want_inf else None
        lam = np.empty((n, n)) if want_inf else None
        L.check(self._lib.gbp_lin_joint_solve(self._h, L.ptr(mu), L.ptr(sigma), L.ptr(eta), L.ptr(lam)))
        return mu, sigma, eta, lam

    def launch_count(self):
        return int(self._lib.gbp_lin_launch_count(self._h))


def device_available():
    """True when the CUDA library is built and sees a device (the generic graph then moves linear pairwise graphs onto it)."""
    try:
        return L.load().gbp_device_count() > 0
    except Exception:
        return False


def tables_from_host_graph(graph):
    """What gbp_lin_create needs, from a graph the client assembled out of hostgraph.VariableNode / Factor objects
    (ndim_posegraph.py:66-87).  Returns None when the graph is not a pairwise linear graph of equal dofs <= 6 with squared loss.
    Pure host code (tested without a GPU)."""
    vs, fs = graph.var_nodes, graph.factors
    if graph.nonlinear_factors or not vs:
        return None
    D = vs[0].dofs
    if D < 1 or D > MAX_DOFS or any(v.dofs != D for v in vs):
        return None
    index = {id(v): k for k, v in enumerate(vs)}
    F = len(fs)
    var_i, var_j = np.zeros(F, np.int32), np.zeros(F, np.int32)
    J, b, var = np.zeros((F, D, 2 * D)), np.zeros((F, D)), np.zeros(F)
    fid = {}
    for k, f in enumerate(fs):
        if len(f.adj_var_nodes) != 2 or f.loss is not None or f.args or any(id(v) not in index for v in f.adj_var_nodes):
            return None
        z = np.atleast_1d(np.asarray(f.measurement, dtype=float))
        if z.ndim != 1 or len(z) > D:
            return None
        x0 = np.asarray(f.linpoint, dtype=float)
        Jf = np.atleast_2d(np.asarray(f.jac_fn(x0), dtype=float))
        h0 = np.atleast_1d(np.asarray(f.meas_fn(x0), dtype=float))
        if Jf.shape != (len(z), 2 * D):
            return None
        var_i[k], var_j[k] = index[id(f.adj_var_nodes[0])], index[id(f.adj_var_nodes[1])]
        J[k, :len(z)] = Jf
        b[k, :len(z)] = Jf @ x0 + z - h0              # gbp/gbp.py:289
        var[k] = f.adaptive_gauss_noise_var
        fid[id(f)] = k
    adj_ptr, adj_msg = np.zeros(len(vs) + 1, np.int32), []
    for k, v in enumerate(vs):
        for f in v.adj_factors:                        # adj_factors order = summation order of update_belief
            if id(f) not in fid:
                return None
            side = [id(a) for a in f.adj_var_nodes].index(id(v))
            adj_msg.append(2 * fid[id(f)] + side)
        adj_ptr[k + 1] = len(adj_msg)
    if len(adj_msg) != 2 * F or len(set(adj_msg)) != 2 * F:
        return None                                    # a factor missing from an adj_factors list (or listed twice)
    prior_eta = np.array([np.asarray(v.prior.eta, dtype=float) for v in vs]).reshape(len(vs), D)
    prior_lam = np.array([np.asarray(v.prior.lam, dtype=float) for v in vs]).reshape(len(vs), D, D)
    msg_eta = np.array([[np.asarray(m.eta, dtype=float) for m in f.messages] for f in fs]).reshape(2 * F, D) if F else np.zeros((0, D))
    msg_lam = np.array([[np.asarray(m.lam, dtype=float) for m in f.messages] for f in fs]).reshape(2 * F, D, D) if F else np.zeros((0, D, D))
    return dict(dim=D, var_i=var_i, var_j=var_j, J=J, b=b, var=var, prior_eta=prior_eta, prior_lam=prior_lam, adj_ptr=adj_ptr,
                adj_msg=np.array(adj_msg, dtype=np.int32), msg_eta=msg_eta, msg_lam=msg_lam)
