"""Bundle-adjustment factor graph on the B200 engine: host-side mirror of the reference's
``gbp/gbp_ba.py`` (BAFactorGraph, Frame/LandmarkVariableNode, ReprojectionFactor,
create_ba_graph) and of the ``FactorGraph`` sweep API of ``gbp/gbp.py:11-153``.

Same names, argument meaning and error behaviour as the reference classes, but no numerics
live here: every method enqueues CUDA work through the C ABI (``gbp_b200._lib``).  Variable
and factor objects are light proxies over lazily refreshed host mirrors of the device
state, so the client code of ``ba.py`` (which loops over ``graph.factors`` reading and
writing ``iters_since_relin``, ``ba.py:91-100``) and of ``vis/ba_vis.py`` (which reads
``node.mu``) runs unmodified with at most one device<->host transfer per field per sweep.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import balio
from .engine import BAEngine, pack_sym, unpack_sym


class _Mirror:
    """Host copy of one device field with valid / dirty tracking."""

    __slots__ = ("field", "data", "valid", "dirty")

    def __init__(self, field):
        self.field, self.data, self.valid, self.dirty = field, None, False, False


class _GaussianView:
    """``NdimGaussian``-shaped view (``.dim``, ``.eta``, ``.lam``; utils/gaussian.py:4-16) of one row
    of a packed [eta | Lambda] field.  Assigning ``.eta`` / ``.lam`` writes through to the device
    (lazily) when the field is writable."""

    __slots__ = ("_g", "_field", "_row", "dim", "_writable")

    def __init__(self, graph, field, row, dim, writable):
        self._g, self._field, self._row, self.dim, self._writable = graph, field, row, dim, writable

    @property
    def eta(self):
        return self._g._mirror(self._field)[self._row, :self.dim].copy()

    @eta.setter
    def eta(self, value):
        self._check()
        self._g._mirror(self._field)[self._row, :self.dim] = np.asarray(value, dtype=np.float64)
        self._g._touch(self._field)

    @property
    def lam(self):
        n = self.dim
        return unpack_sym(self._g._mirror(self._field)[self._row, n:n + n * (n + 1) // 2], n)

    @lam.setter
    def lam(self, value):
        self._check()
        n = self.dim
        value = np.asarray(value, dtype=np.float64)
        if value.shape != (n, n):
            raise ValueError(f"lam must be {n}x{n}")
        self._g._mirror(self._field)[self._row, n:n + n * (n + 1) // 2] = pack_sym(value)
        self._g._touch(self._field)

    def _check(self):
        if not self._writable:
            raise AttributeError("this Gaussian is a read-only view of device state")


class _VariableNode:
    """VariableNode proxy (gbp/gbp.py:156-198)."""

    _is_cam = True

    def __init__(self, graph, index):
        self._g, self._i = graph, index

    @property
    def dofs(self):
        return 6 if self._is_cam else 3

    @property
    def variableID(self):
        return self._i if self._is_cam else self._g._eng.C + self._i

    def _bfield(self):
        return L.F_CAM_BELIEF if self._is_cam else L.F_LMK_BELIEF

    @property
    def mu(self):
        n = self.dofs
        g = self._g
        bm = g._mirrors.get(self._bfield())
        if bm is not None and bm.valid and not g._snap_pending:
            return bm.data[self._i, -n:].copy()            # full belief table already on the host (maybe edited)
        return g._mirror(L.F_CAM_MU if self._is_cam else L.F_LMK_MU)[self._i].copy()

    @mu.setter
    def mu(self, value):
        n = self.dofs
        self._g._mirror(self._bfield())[self._i, -n:] = np.asarray(value, dtype=np.float64)
        self._g._touch(self._bfield())

    @property
    def belief(self):
        return _GaussianView(self._g, self._bfield(), self._i, self.dofs, True)

    @property
    def prior(self):
        return _GaussianView(self._g, L.F_CAM_PRIOR if self._is_cam else L.F_LMK_PRIOR, self._i, self.dofs, True)

    @property
    def Sigma(self):
        return np.linalg.inv(self.belief.lam)      # presentation only; the sweep never needs it

    @property
    def adj_factors(self):
        return [self._g.factors[int(f)] for f in self._g._adjacent_factors(self._is_cam, self._i)]

    def update_belief(self):
        raise NotImplementedError("beliefs are updated for the whole graph on the device: "
                                  "call graph.update_all_beliefs()")


class FrameVariableNode(_VariableNode):
    """gbp/gbp_ba.py:78-81"""
    _is_cam = True

    @property
    def c_id(self):
        return self._i


class LandmarkVariableNode(_VariableNode):
    """gbp/gbp_ba.py:72-75"""
    _is_cam = False

    @property
    def l_id(self):
        return self._i


class ReprojectionFactor:
    """Factor proxy (gbp/gbp.py:201-373, gbp/gbp_ba.py:84-94)."""

    def __init__(self, graph, index):
        self._g, self._f = graph, index

    factorID = property(lambda self: self._f)

    @property
    def adj_vIDs(self):
        c, l = self._g._adj()[self._f]
        return [int(c), int(self._g._eng.C + l)]

    @property
    def adj_var_nodes(self):
        c, l = self._g._adj()[self._f]
        return [self._g.cam_nodes[int(c)], self._g.lmk_nodes[int(l)]]

    @property
    def adj_beliefs(self):
        c, l = self._g._adj()[self._f]
        return [_GaussianView(self._g, L.F_CAM_BELIEF, int(c), 6, False),
                _GaussianView(self._g, L.F_LMK_BELIEF, int(l), 3, False)]

    @property
    def messages(self):
        return [_GaussianView(self._g, L.F_MSG_CAM, self._f, 6, True),
                _GaussianView(self._g, L.F_MSG_LMK, self._f, 3, True)]

    @property
    def dofs_conditional_vars(self):
        return 9

    @property
    def linpoint(self):
        return self._g._mirror(L.F_LINPOINT)[self._f].copy()

    @property
    def measurement(self):
        return self._g._mirror(L.F_MEASUREMENT)[self._f].copy()

    @property
    def iters_since_relin(self):
        return int(self._g._mirror(L.F_ITERS)[self._f, 0])

    @iters_since_relin.setter
    def iters_since_relin(self, value):
        self._g._mirror(L.F_ITERS)[self._f, 0] = int(value)
        self._g._touch(L.F_ITERS)

    @property
    def eta_damping(self):
        return self._g.eta_damping if (int(self._g._mirror(L.F_FLAGS)[self._f, 0]) & 1) else 0.0

    @eta_damping.setter
    def eta_damping(self, value):
        # the engine stores per-factor damping as one bit: on (= graph.eta_damping) or off
        if value == 0.0:
            on = 0
        elif value == self._g.eta_damping:
            on = 1
        else:
            raise ValueError("per-factor eta_damping must be 0 or graph.eta_damping on the B200 engine")
        m = self._g._mirror(L.F_FLAGS)
        m[self._f, 0] = (int(m[self._f, 0]) & ~1) | on
        self._g._touch(L.F_FLAGS)

    @property
    def robust_flag(self):
        return bool(int(self._g._mirror(L.F_FLAGS)[self._f, 0]) & 2)

    @property
    def gauss_noise_var(self):
        return self._g._var0

    @property
    def adaptive_gauss_noise_var(self):
        if self._g.loss is None:
            return self._g._var0
        return float(self._g._mirror(L.F_ADAPTIVE_VAR)[self._f, 0])

    @property
    def loss(self):
        return self._g.loss

    @property
    def mahalanobis_threshold(self):
        return self._g.Nstds

    @property
    def args(self):
        return (self._g.K,)

    @property
    def factor(self):
        """(eta_f, Lambda_f) = (J^T b, J^T J)/var at the current linearisation point (gbp/gbp.py:287-289)."""
        jb = self._g._mirror(L.F_JACOBIAN_B)[self._f]
        J, b = jb[:18].reshape(2, 9), jb[18:]
        var = self.adaptive_gauss_noise_var

        class _G:
            pass
        g = _G()
        g.dim, g.eta, g.lam = 9, J.T @ b / var, J.T @ J / var
        return g

    def compute_residual(self):
        """gbp/gbp.py:251-259 (evaluated on the device for the whole graph, cached per sweep)."""
        return self._g._residuals()[self._f].copy()

    def reprojection_err(self):
        """gbp/gbp_ba.py:90-94"""
        return float(np.linalg.norm(self.compute_residual()))

    def energy(self):
        """gbp/gbp.py:261-265"""
        return 0.5 * np.linalg.norm(self.compute_residual()) ** 2 / self.adaptive_gauss_noise_var


class _LazySeq:
    """Sized, iterable, indexable sequence that builds proxy objects on demand (a 10 M-factor
    graph cannot afford one Python object per factor)."""

    def __init__(self, n, make):
        self._n, self._make = n, make

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._make(j) for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._make(i)

    def __iter__(self):
        for i in range(self._n):
            yield self._make(i)

    def __add__(self, other):
        return _ConcatSeq(self, other)


class _ConcatSeq(_LazySeq):
    def __init__(self, a, b):
        self._a, self._b = a, b
        super().__init__(len(a) + len(b), lambda i: a[i] if i < len(a) else b[i - len(a)])


_FACTOR_FIELDS_INVALIDATED_BY_SWEEP = (L.F_MSG_CAM, L.F_MSG_LMK, L.F_LINPOINT, L.F_ITERS, L.F_FLAGS,
                                       L.F_ADAPTIVE_VAR, L.F_JACOBIAN_B)
_BELIEF_FIELDS = (L.F_CAM_BELIEF, L.F_LMK_BELIEF, L.F_CAM_MU, L.F_LMK_MU)
_MU_FIELDS = (L.F_CAM_MU, L.F_LMK_MU)


class BAFactorGraph:
    """gbp/gbp_ba.py:12-69 on top of gbp/gbp.py:11-153, device resident."""

    def __init__(self, problem: balio.BALProblem, configs: dict, device=0, stream=None, tile_edges=0, lmk_block=0,
                 kernel_variant=0, chunks=None):
        self.nonlinear_factors = True
        self._eta_damping = float(configs["eta_damping"])
        self._beta = float(configs["beta"])
        self._num_undamped_iters = int(configs["num_undamped_iters"])
        self._min_linear_iters = int(configs["min_linear_iters"])
        self.loss = configs.get("loss", None)
        self.Nstds = float(configs.get("Nstds", 3.0))
        self._var0 = float(configs["gauss_noise_std"]) ** 2
        self.K = problem.K
        self._eng = BAEngine(problem.cam_id, problem.lmk_id, problem.z, problem.cam_means, problem.lmk_means,
                             problem.K4, configs, device=device, stream=stream, tile_edges=tile_edges,
                             lmk_block=lmk_block, kernel_variant=kernel_variant, chunks=chunks)
        e = self._eng
        self.cam_nodes = _LazySeq(e.C, lambda i: FrameVariableNode(self, i))
        self.lmk_nodes = _LazySeq(e.L, lambda i: LandmarkVariableNode(self, i))
        self.var_nodes = self.cam_nodes + self.lmk_nodes
        self.factors = _LazySeq(e.F, lambda i: ReprojectionFactor(self, i))
        self.n_var_nodes = e.C + e.L
        self.n_factor_nodes = e.F
        self.n_edges = 2 * e.F
        self._mirrors = {}
        self._adj_cache = None
        self._csr = {}
        self._res_cache = None
        self._metric_cache = None
        self._params_dirty = False
        # small graphs: metrics + belief tables are copied to pinned host buffers right behind every sweep
        self._eager = (e.C * 6 + e.L * 3) * 8 <= self._EAGER_MAX_BYTES and e.F > 0
        self._snap_pending = False      # an asynchronous snapshot is in flight / unread
        self._snap_metrics_ok = False   # ... and its metrics still describe the device state
        self._snap_region = None        # pinned host image of [metrics | keyframe beliefs | landmark beliefs]
        self._snap_metrics = None

    # ------------------------------------------------------------------ FactorGraph parameters
    def _param(name):  # noqa: N805
        def get(self):
            return getattr(self, "_" + name)

        def set_(self, v):
            setattr(self, "_" + name, v)
            self._params_dirty = True
        return property(get, set_)

    eta_damping = _param("eta_damping")
    beta = _param("beta")
    num_undamped_iters = _param("num_undamped_iters")
    min_linear_iters = _param("min_linear_iters")
    del _param

    # ------------------------------------------------------------------ mirrors
    def _resolve_pending(self):
        """Wait for the snapshot enqueued behind the last sweep; its belief tables become the valid mirrors."""
        if not self._snap_pending:
            return
        self._eng.snapshot_wait()
        self._snap_pending = False
        for f in _MU_FIELDS:
            m = self._mirrors[f]
            m.valid, m.dirty = True, False
        if self._snap_metrics_ok:
            a, en, n = (float(v) for v in self._snap_metrics)
            self._metric_cache = (a / self._eng.F, en, int(round(n)))

    def _snapshot_region(self):
        """Pinned host image of the device's snapshot region [metrics | keyframe means | landmark means]; the mean
        mirrors are views into it."""
        if self._snap_region is None:
            e = self._eng
            total, o_met, o_cam, o_lmk = e.snapshot_layout()
            reg = L.pinned_empty((total,), np.uint8)
            self._snap_region = reg
            self._snap_metrics = reg[o_met:o_met + 24].view(np.float64)
            for f, off, rows, w in ((L.F_CAM_MU, o_cam, e.C, 6), (L.F_LMK_MU, o_lmk, e.L, 3)):
                m = self._mirrors.setdefault(f, _Mirror(f))
                view = reg[off:off + rows * w * 8].view(np.float64).reshape(rows, w)
                if m.data is not None and m.valid:
                    view[:] = m.data
                m.data = view
        return self._snap_region

    def _enqueue_snapshot(self):
        self._eng.snapshot_async(self._snapshot_region())
        self._snap_pending, self._snap_metrics_ok = True, True

    def _mirror(self, field):
        if self._snap_pending and field in _BELIEF_FIELDS:
            self._resolve_pending()
        if field in _MU_FIELDS:
            # a belief table edited on the host (client assigned node.mu / belief) supersedes the compact means
            bm = self._mirrors.get(L.F_CAM_BELIEF if field == L.F_CAM_MU else L.F_LMK_BELIEF)
            if bm is not None and bm.dirty:
                self._flush()
        m = self._mirrors.get(field)
        if m is None:
            m = self._mirrors[field] = _Mirror(field)
        if not m.valid:
            m.data = self._eng.read(field, m.data)
            m.valid, m.dirty = True, False
        return m.data

    def _touch(self, field):
        self._mirrors[field].dirty = True
        self._res_cache = None
        self._metric_cache = None
        self._snap_metrics_ok = False
        if field == L.F_LINPOINT:
            jm = self._mirrors.get(L.F_JACOBIAN_B)
            if jm is not None:
                jm.valid = False

    def _flush(self):
        """Push host-side edits (client writes through the proxies) to the device."""
        if self._params_dirty:
            self._eng.set_params(self._eta_damping, self._beta, self._num_undamped_iters, self._min_linear_iters)
            self._params_dirty = False
        for m in self._mirrors.values():
            if m.dirty:
                if m.field == L.F_ITERS and m.data.size and (m.data == m.data.flat[0]).all():
                    self._eng.fill_iters(int(m.data.flat[0]))   # the loop of ba.py:91-93
                else:
                    self._eng.write(m.field, m.data)
                m.dirty = False
                if m.field in (L.F_CAM_BELIEF, L.F_LMK_BELIEF):
                    # the device refreshed its compact means from the table just written; the host copy is stale
                    mm = self._mirrors.get(L.F_CAM_MU if m.field == L.F_CAM_BELIEF else L.F_LMK_MU)
                    if mm is not None:
                        mm.valid = False

    def _invalidate(self, fields):
        self._snap_metrics_ok = False
        if self._snap_pending and any(f in _BELIEF_FIELDS for f in fields):
            # the in-flight copy targets the mirrors' buffers: let it land before they are reused
            self._eng.snapshot_wait()
            self._snap_pending = False
        for f in fields:
            m = self._mirrors.get(f)
            if m is not None:
                m.valid = False
        self._res_cache = None
        self._metric_cache = None

    def _adj(self):
        if self._adj_cache is None:
            self._adj_cache = self._eng.read(L.F_ADJ)
        return self._adj_cache

    def _adjacent_factors(self, is_cam, i):
        key = "cam" if is_cam else "lmk"
        if key not in self._csr:
            ids = self._adj()[:, 0 if is_cam else 1]
            order = np.argsort(ids, kind="stable")
            n = self._eng.C if is_cam else self._eng.L
            ptr = np.searchsorted(ids[order], np.arange(n + 1))
            self._csr[key] = (order, ptr)
        order, ptr = self._csr[key]
        return order[ptr[i]:ptr[i + 1]]

    def _residuals(self):
        """h(mu_cam, mu_lmk) - z of every factor, from the device-side model at the current means."""
        if self._res_cache is None:
            from .engine import reprojection_eval
            adj = self._adj()
            cm = self._mirror(L.F_CAM_MU)
            lm = self._mirror(L.F_LMK_MU)
            x = np.concatenate([cm[adj[:, 0]], lm[adj[:, 1]]], axis=1)
            h, _ = reprojection_eval(x, self._eng.K4, self._eng.device)
            self._res_cache = h - self._mirror(L.F_MEASUREMENT)
        return self._res_cache

    # ------------------------------------------------------------------ BAFactorGraph API
    def generate_priors_var(self, weaker_factor=100):
        """gbp/gbp_ba.py:20-34"""
        self._flush()
        self._eng.generate_priors(weaker_factor)
        self._invalidate((L.F_CAM_PRIOR, L.F_LMK_PRIOR))

    def weaken_priors(self, weakening_factor):
        """gbp/gbp_ba.py:36-42"""
        self._flush()
        self._eng.scale_priors(weakening_factor)
        self._invalidate((L.F_CAM_PRIOR, L.F_LMK_PRIOR))

    def set_priors_var(self, priors):
        """gbp/gbp_ba.py:44-52: `priors` = one covariance matrix per variable node (cameras first)."""
        e = self._eng
        if len(priors) != e.C + e.L:
            raise ValueError("set_priors_var needs one covariance per variable node")
        self._flush()
        cam = np.stack([np.linalg.inv(np.asarray(p, dtype=np.float64)) for p in priors[:e.C]]) if e.C else np.zeros((0, 6, 6))
        lmk = np.stack([np.linalg.inv(np.asarray(p, dtype=np.float64)) for p in priors[e.C:]]) if e.L else np.zeros((0, 3, 3))
        e.set_priors(pack_sym(cam), pack_sym(lmk))
        self._invalidate((L.F_CAM_PRIOR, L.F_LMK_PRIOR))

    def compute_residuals(self):
        """gbp/gbp_ba.py:54-59"""
        return list(self._residuals().ravel())

    # Small graphs: the first look at the state after a sweep (are / energy / a mean) fetches the metrics
    # AND both belief tables with one stream synchronisation; larger graphs fetch lazily per field.
    _EAGER_MAX_BYTES = 1 << 20

    def _snapshot(self):
        if self._snap_pending and self._snap_metrics_ok:
            self._resolve_pending()
        if self._metric_cache is not None:
            return self._metric_cache
        self._resolve_pending()
        self._flush()
        e = self._eng
        if self._eager:
            # fetch metrics and both belief tables with one copy and one synchronisation
            self._eng.snapshot_async(self._snapshot_region())
            self._snap_pending, self._snap_metrics_ok = True, True
            self._resolve_pending()
            return self._metric_cache
        a, en, n = e.metrics()
        self._metric_cache = (a / e.F if e.F else 0.0, en, n)
        return self._metric_cache

    def are(self):
        """gbp/gbp_ba.py:61-69"""
        return self._snapshot()[0]

    def energy(self):
        """gbp/gbp.py:36-44"""
        return self._snapshot()[1]

    def metrics(self):
        """ARE, energy and the count of `iters_since_relin == 0` (ba.py:95-100) in ONE device pass."""
        return self._snapshot()

    def n_relinearising(self):
        return self.metrics()[2]

    # ------------------------------------------------------------------ FactorGraph sweep API
    def _sweep(self, stages):
        self._flush()
        e = self._eng
        e.sweep_local(stages)
        if stages & L.ST_BELIEFS:
            e.cam_update()
            self._invalidate(_BELIEF_FIELDS)
        self._invalidate(_FACTOR_FIELDS_INVALIDATED_BY_SWEEP)

    def synchronous_iteration(self, local_relin=True, robustify=False):
        """gbp/gbp.py:86-92"""
        self._flush()
        if self._eager:
            # sweep + beliefs + metrics + copies to the pinned mirrors: one CUDA-graph launch
            self._invalidate(_BELIEF_FIELDS + _FACTOR_FIELDS_INVALIDATED_BY_SWEEP)
            self._eng.iterate_snapshot(robustify, local_relin, self._snapshot_region())
            self._snap_pending, self._snap_metrics_ok = True, True
            return
        self._eng.iterate(1, robustify=robustify, local_relin=local_relin)
        self._invalidate(_BELIEF_FIELDS + _FACTOR_FIELDS_INVALIDATED_BY_SWEEP)

    def iterate(self, n_iters, local_relin=True, robustify=False):
        """n x synchronous_iteration without returning to Python in between (engine extension)."""
        self._flush()
        self._eng.iterate(n_iters, robustify=robustify, local_relin=local_relin)
        self._invalidate(_BELIEF_FIELDS + _FACTOR_FIELDS_INVALIDATED_BY_SWEEP)
        if self._eager:
            self._enqueue_snapshot()

    def reset_iters_since_relin(self, value=1):
        """The client loop of ba.py:91-93 (`for factor in graph.factors: factor.iters_since_relin = 1`) as one call
        (engine extension; assigning through the factor proxies does the same with one device fill)."""
        self._flush()
        self._eng.fill_iters(int(value))
        self._invalidate((L.F_ITERS,))

    def robustify_all_factors(self):
        """gbp/gbp.py:82-84"""
        self._sweep(L.ST_ROBUSTIFY)

    def relinearise_factors(self):
        """gbp/gbp.py:64-80"""
        self._sweep(L.ST_RELIN)

    def compute_all_messages(self, local_relin=True):
        """gbp/gbp.py:46-54"""
        self._sweep(L.ST_MESSAGES | (L.ST_LOCAL_DAMPING if local_relin else 0))

    def update_all_beliefs(self):
        """gbp/gbp.py:56-58"""
        self._flush()
        self._eng.update_beliefs()
        self._invalidate(_BELIEF_FIELDS)

    def compute_all_factors(self):
        """gbp/gbp.py:60-62: relinearise every factor at the current adjacent belief means."""
        adj = self._adj()
        cm = self._mirror(L.F_CAM_MU)
        lm = self._mirror(L.F_LMK_MU)
        self._mirror(L.F_LINPOINT)[:] = np.concatenate([cm[adj[:, 0]], lm[adj[:, 1]]], axis=1)
        self._touch(L.F_LINPOINT)
        self._flush()

    def get_means(self):
        """gbp/gbp.py:146-153"""
        return np.concatenate([self._mirror(L.F_CAM_MU).ravel(), self._mirror(L.F_LMK_MU).ravel()])

    def joint_distribution_inf(self):
        """gbp/gbp.py:94-134 (dense; small graphs only)."""
        e = self._eng
        n = 6 * e.C + 3 * e.L
        if n > 20000:
            raise MemoryError("joint_distribution_inf builds a dense matrix; graph too large")
        eta, lam = np.zeros(n), np.zeros((n, n))
        cp, lp = self._mirror(L.F_CAM_PRIOR), self._mirror(L.F_LMK_PRIOR)
        for c in range(e.C):
            eta[6 * c:6 * c + 6] = cp[c, :6]
            lam[6 * c:6 * c + 6, 6 * c:6 * c + 6] = unpack_sym(cp[c, 6:], 6)
        o = 6 * e.C
        for l in range(e.L):
            eta[o + 3 * l:o + 3 * l + 3] = lp[l, :3]
            lam[o + 3 * l:o + 3 * l + 3, o + 3 * l:o + 3 * l + 3] = unpack_sym(lp[l, 3:], 3)
        jb = self._mirror(L.F_JACOBIAN_B)
        adj = self._adj()
        var = (np.full(e.F, self._var0) if self.loss is None else self._mirror(L.F_ADAPTIVE_VAR)[:, 0])
        for f in range(e.F):
            J, b = jb[f, :18].reshape(2, 9), jb[f, 18:]
            ix = np.concatenate([np.arange(6) + 6 * adj[f, 0], np.arange(3) + o + 3 * adj[f, 1]])
            eta[ix] += J.T @ b / var[f]
            lam[np.ix_(ix, ix)] += J.T @ J / var[f]
        return eta, lam

    def joint_distribution_cov(self):
        """gbp/gbp.py:136-144"""
        eta, lam = self.joint_distribution_inf()
        sigma = np.linalg.inv(lam)
        return sigma @ eta, sigma

    def reset(self):
        """Engine extension: back to the state right after create_ba_graph."""
        self._resolve_pending()
        self._flush()
        self._eng.reset()
        for m in self._mirrors.values():
            m.valid = False
        self._res_cache = None
        self._metric_cache = None

    def close(self):
        if self._snap_pending:
            self._eng.snapshot_wait()
            self._snap_pending = False
        self._eng.close()


def create_ba_graph(bal_file, configs, device=0, stream=None, tile_edges=0, lmk_block=0, kernel_variant=0, chunks=None):
    """gbp/gbp_ba.py:97-150: build the graph object from a BAL-style file (text or .npz)."""
    problem = bal_file if isinstance(bal_file, balio.BALProblem) else balio.read_bal(bal_file)
    return BAFactorGraph(problem, configs, device=device, stream=stream, tile_edges=tile_edges, lmk_block=lmk_block,
                         kernel_variant=kernel_variant, chunks=chunks)
