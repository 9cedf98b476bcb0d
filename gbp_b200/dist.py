"""Multi-GPU GBP sweep: the factor graph is cut by LANDMARK across the ranks of one box.

Every factor lives with its landmark, so all landmarks are interior (their incoming messages
are local) and the boundary variables are the keyframes, replicated on every rank.  Per
synchronous iteration each rank runs the sweep over its own edges, reduces its local
factor->keyframe messages to one partial (eta, Lambda) sum per keyframe (C x 27 doubles), and the
ranks exchange those partial sums with ONE all-gather (NCCL over NVLink / NVSwitch; gloo in the
CPU tests); every rank then adds prior + the partials in rank order, so the keyframe beliefs are
bit-identical everywhere.  ARE / energy need a 3-scalar all-reduce only when the client asks.

`p2p=True`: the exchange without a collective call.  Every rank writes its partial sums straight into the other
ranks' exchange buffers over NVLink (CUDA IPC mappings of one small buffer per rank) and raises per-CTA flags
there; the keyframe update kernel waits on the flags of its own buffer (`gbp_ba_p2p_*`, kernels
`p2p_scatter_kernel` / `p2p_gather_update_kernel`).  Same rank-ordered sum, so the same bits.

Streams: with world > 1 the engine's kernels and the collective must be ordered on ONE stream.  The graph takes a
single `torch_stream` (a torch.cuda.Stream; created here when omitted), hands its raw handle to the engine and
issues the collective under `torch.cuda.stream(torch_stream)`.

The reference has no distributed code; this is new functionality behind the same
`synchronous_iteration` surface.

The compute engine is injectable (`engine_factory`) so that the partition / exchange / merge logic
is testable on CPU with world_size 2 (tests/test_dist_cpu.py); the product always uses the CUDA
engine.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .balio import BALProblem


def landmark_partition(n_lmks: int, world: int):
    """Contiguous landmark blocks: rank r owns [bounds[r], bounds[r+1])."""
    return [(n_lmks * r) // world for r in range(world + 1)]


def local_problem(prob: BALProblem, rank: int, world: int):
    """The sub-problem of one rank: all keyframes, its landmark block, the measurements of those landmarks
    (file order preserved).  Returns (BALProblem with local landmark ids, global index of each local measurement)."""
    b = landmark_partition(prob.n_points, world)
    l0, l1 = b[rank], b[rank + 1]
    sel = np.nonzero((prob.lmk_id >= l0) & (prob.lmk_id < l1))[0]
    sub = BALProblem(prob.cam_id[sel], prob.lmk_id[sel] - l0, prob.z[sel], prob.cam_means, prob.lmk_means[l0:l1], prob.K4)
    return sub, sel, (l0, l1)


class _DevArray:
    """Zero-copy view of engine-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class CudaEngineAdapter:
    """The CUDA engine seen through the four operations the exchange layer needs."""

    def __init__(self, sub: BALProblem, configs, device, stream, **kw):
        import torch
        from .engine import BAEngine
        self.eng = BAEngine(sub.cam_id, sub.lmk_id, sub.z, sub.cam_means, sub.lmk_means, sub.K4, configs, device=device,
                            stream=stream, **kw)
        ptr, nbytes = self.eng.device_ptr(L.F_CAM_PARTIAL)
        self._partial = torch.as_tensor(_DevArray(ptr, nbytes // 8), device=f"cuda:{device}")
        self._torch = torch

    C = property(lambda self: self.eng.C)
    L = property(lambda self: self.eng.L)
    F = property(lambda self: self.eng.F)

    def prior_scan(self):
        return self._torch.from_numpy(self.eng.prior_scan())

    def generate_priors(self, weaker, cam_max):
        self.eng.generate_priors(weaker, cam_max.cpu().numpy())

    def scale_priors(self, f):
        self.eng.scale_priors(f)

    def sweep_local(self, stages):
        self.eng.sweep_local(stages)

    def landmark_update(self):
        self.eng.landmark_update()

    def partial_tensor(self):
        return self._partial

    def new_gather_buffer(self, world):
        return self._torch.empty(world * self._partial.numel(), dtype=self._torch.float64, device=self._partial.device)

    def apply_gathered(self, gathered, world):
        self.eng.cam_update(gathered.data_ptr(), world)

    # ---- peer-memory exchange (opt-in)
    def p2p_setup(self, rank, world, dist):
        handle = self.eng.p2p_init(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, handle)        # 64-byte CUDA IPC handles, rank order
        self.eng.p2p_attach(handles)
        dist.barrier()                                  # everybody mapped everybody before the first store

    def p2p_scatter(self):
        self.eng.p2p_scatter()

    def p2p_gather_update(self):
        self.eng.p2p_gather_update()

    def p2p_status(self):
        return self.eng.p2p_status()

    def iterate_single(self, robustify, local_relin):
        self.eng.iterate(1, robustify=robustify, local_relin=local_relin)

    def update_beliefs_single(self):
        self.eng.update_beliefs()

    def metrics(self):
        a, e, n = self.eng.metrics()
        return np.array([a, e, float(n)])

    def cam_means(self):
        return self.eng.read(L.F_CAM_MU)

    def lmk_means(self):
        return self.eng.read(L.F_LMK_MU)

    def fill_iters(self, v):
        self.eng.fill_iters(v)

    def reset(self):
        self.eng.reset()

    def close(self):
        self.eng.close()


class PartitionedBAGraph:
    """`synchronous_iteration` / `generate_priors_var` / `are` / `energy` over a landmark-partitioned graph."""

    def __init__(self, prob: BALProblem, configs, rank=0, world=1, device=0, stream=None, dist=None,
                 engine_factory=None, torch_stream=None, p2p=False, **engine_kw):
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs an initialised torch.distributed module")
        if engine_factory is None and world > 1:
            # ONE stream for the engine's kernels and the collective: the raw handle is derived from the torch stream
            import torch
            if torch_stream is None:
                if stream is not None:
                    raise ValueError("world > 1: pass torch_stream (a torch.cuda.Stream), not a raw stream handle; the "
                                     "collective has to be ordered on the engine's stream")
                torch_stream = torch.cuda.Stream(device=device)
            if stream is not None and int(stream) != int(torch_stream.cuda_stream):
                raise ValueError("stream and torch_stream name different CUDA streams")
            stream = torch_stream.cuda_stream
        engine_kw_stream = torch_stream
        self.rank, self.world, self.dist = rank, world, dist
        self.n_iterations = 0        # synchronous iterations applied to the state since creation / reset
        self.F_total, self.L_total, self.C = prob.n_edges, prob.n_points, prob.n_keyframes
        sub, self.local_measurements, self.lmk_range = local_problem(prob, rank, world)
        factory = engine_factory or (lambda s, c: CudaEngineAdapter(s, c, device, stream, **engine_kw))
        self.adapter = factory(sub, configs)
        self._gather = self.adapter.new_gather_buffer(world) if world > 1 else None
        self.p2p = bool(p2p) and world > 1
        if self.p2p:
            if not hasattr(self.adapter, "p2p_setup"):
                raise ValueError("this engine has no peer-memory exchange (p2p=True needs the CUDA engine)")
            self.adapter.p2p_setup(rank, world, dist)
        self._graphs = {}        # stages -> captured CUDA graph of [local sweep, all-gather, keyframe update]
        self._torch_stream = engine_kw_stream

    @property
    def engine(self):
        return self.adapter.eng

    # ------------------------------------------------------------------ exchange
    def _exchange_and_update(self):
        """keyframe partial sums -> all ranks (one all-gather); the landmark beliefs, which need no communication, are
        updated on the compute stream while the collective is in flight; then prior + partials in rank order."""
        a = self.adapter
        if self.p2p:
            a.p2p_scatter()             # partial sums -> every peer's buffer (stores over NVLink) + flags
            a.landmark_update()         # overlaps the transfer
            a.p2p_gather_update()       # waits on this rank's flags; prior + sums in rank order
            return
        with self._stream_ctx():
            work = self.dist.all_gather_into_tensor(self._gather, a.partial_tensor(), async_op=True)
            a.landmark_update()
            work.wait()
            a.apply_gathered(self._gather, self.world)

    def _stream_ctx(self):
        """The collective is issued (and waited for) on the engine's stream."""
        if self._torch_stream is None:
            import contextlib
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(self._torch_stream)

    # ------------------------------------------------------------------ API
    def generate_priors_var(self, weaker_factor=100):
        """gbp/gbp_ba.py:20-34 with the per-keyframe maximum taken over all ranks."""
        a = self.adapter
        if self.world == 1:
            a.generate_priors(weaker_factor, a.prior_scan())
            return
        cam_max = a.prior_scan()
        dev = self._gather.device
        t = cam_max.to(dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        a.generate_priors(weaker_factor, t.cpu())

    def weaken_priors(self, f):
        self.adapter.scale_priors(f)

    def update_all_beliefs(self):
        if self.world == 1:
            self.adapter.update_beliefs_single()
        else:
            self.adapter.sweep_local(L.ST_BELIEFS | L.ST_DEFER_LANDMARKS)
            self._exchange_and_update()

    def synchronous_iteration(self, local_relin=True, robustify=False):
        """gbp/gbp.py:86-92 over the partitioned graph: local sweep -> one all-gather -> keyframe beliefs."""
        self.n_iterations += 1
        if self.world == 1:
            self.adapter.iterate_single(robustify, local_relin)
            return
        st = L.ST_MESSAGES | L.ST_BELIEFS | L.ST_DEFER_LANDMARKS
        if robustify:
            st |= L.ST_ROBUSTIFY
        if local_relin:
            st |= L.ST_RELIN | L.ST_LOCAL_DAMPING
        g = self._graphs.get(st)
        if g is not None:
            g.replay()
            return
        self.adapter.sweep_local(st)
        self._exchange_and_update()

    def capture(self, local_relin=True, robustify=False):
        """Capture [local sweep -> all-gather -> keyframe update] of one synchronous iteration into a CUDA graph
        (NCCL collectives are capturable), so that an iteration is ONE launch per rank instead of three engine
        calls plus a Python-side collective.  Does NOT advance the state (no iteration is applied): N calls of
        synchronous_iteration apply N iterations whether or not capture() was called in between."""
        if self.world == 1 or self._torch_stream is None:
            return False
        import torch
        st = L.ST_MESSAGES | L.ST_BELIEFS | L.ST_DEFER_LANDMARKS | (L.ST_ROBUSTIFY if robustify else 0) | \
            ((L.ST_RELIN | L.ST_LOCAL_DAMPING) if local_relin else 0)
        if st in self._graphs:
            return True
        # NCCL sets up its channels on the first collective, which must happen outside the capture: gather the current
        # partial sums into the scratch buffer once (touches no state of the solve)
        if not self.p2p:
            with self._stream_ctx():
                self.dist.all_gather_into_tensor(self._gather, self.adapter.partial_tensor())
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._torch_stream, capture_error_mode="thread_local"):
            self.adapter.sweep_local(st)
            self._exchange_and_update()
        self._graphs[st] = g
        return True

    def fill_iters(self, value):
        self.adapter.fill_iters(value)

    def reset(self):
        """Back to the state right after construction (gbp_ba_reset on every rank): zero messages and priors."""
        self.adapter.reset()
        self.n_iterations = 0

    def metrics(self):
        """(ARE, energy, number of factors with iters_since_relin == 0) over the WHOLE graph."""
        self._check_exchange()
        m = self.adapter.metrics()
        if self.world > 1:
            import torch
            t = torch.from_numpy(m).to(self._gather.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            m = t.cpu().numpy()
        return float(m[0]) / self.F_total, float(m[1]), int(round(float(m[2])))

    def are(self):
        return self.metrics()[0]

    def energy(self):
        return self.metrics()[1]

    def _check_exchange(self):
        """A peer-memory exchange that timed out (a peer died or never launched) poisons the keyframe beliefs with NaN on the
        device; turn it into an exception as soon as the client looks at results."""
        if self.p2p:
            done, timeouts = self.adapter.p2p_status()
            if timeouts:
                raise RuntimeError(f"rank {self.rank}: {timeouts} peer-memory exchange wait(s) timed out after {done} exchanges; "
                                   "the keyframe beliefs are invalid")

    def get_means(self):
        """All belief means in variable order (keyframes, then landmarks) on every rank."""
        self._check_exchange()
        cam = self.adapter.cam_means().ravel()
        lmk = self.adapter.lmk_means()
        if self.world > 1:
            parts = [None] * self.world
            self.dist.all_gather_object(parts, lmk)
            lmk = np.concatenate(parts, axis=0)
        return np.concatenate([cam, lmk.ravel()])

    def close(self):
        # captured graphs hold NCCL work: they must die before the process group is destroyed
        if self._graphs:
            import torch
            torch.cuda.synchronize()
            self._graphs.clear()
        if self.p2p:
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()          # nobody still writes into a buffer that is about to be freed
        self.adapter.close()
