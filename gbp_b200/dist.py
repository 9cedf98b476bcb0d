"""Multi-GPU GBP sweep: the factor graph is cut by LANDMARK across the ranks of one box.

Every factor lives with its landmark, so all landmarks are interior (their incoming messages
are local) and the boundary variables are the keyframes, replicated on every rank.  Per
synchronous iteration each rank runs the sweep over its own edges, reduces its local
factor->keyframe messages to one partial (eta, Lambda) sum per keyframe (C x 27 doubles), and the
ranks exchange those partial sums with ONE all-gather (NCCL over NVLink / NVSwitch; gloo in the
CPU tests); every rank then adds the partials in rank order and the prior, so the keyframe beliefs are
bit-identical everywhere.  ARE / energy need a 3-scalar all-reduce only when the client asks.

Bit-identical across the NUMBER of GPUs as well: the engine forms the keyframe-side sums per landmark CHUNK (gbp_config:
8 chunks of consecutive landmarks from 65536 landmarks on) and adds the chunk sums in chunk order; a rank holds whole chunks of
that global chunking (its partial "sum" is its chunk sums, K / world x C x 27 doubles) and lays out exactly the tiles the
single-GPU plan has for them (tile size, landmark blocks and kernel build are chosen from the GLOBAL sizes).  So 1, 2, 4 and 8
GPUs run the same floating-point operations in the same order and produce the same bits.

(A collective-free exchange -- every rank storing its partial sums into the other ranks' buffers over NVLink through CUDA IPC
mappings, per-CTA flags -- was built and run on 2 and 8 GPUs in round 2: same bits, same speed as the all-gather within 1 %
(0.2143 vs 0.2130 ms per iteration at 8 GPUs), so it was removed again; profiles/r2e_bench_n8_*.json.)

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


def global_layout(prob: BALProblem, world: int):
    """Engine layout choices made from the GLOBAL sizes, so that every rank lays out its share exactly like the single-GPU plan
    (the rules of choose_tiling / chunk_bounds_of / the streaming switch in gbp_ba.cu): tile size, landmark block, kernel build,
    number of landmark chunks (a multiple of world)."""
    F, Lm = prob.n_edges, prob.n_points
    T = 64 if F >= 64 * 148 * 6 else 32
    lblock = max(Lm, 1) if Lm * 96 <= (24 << 20) else 262144
    k_auto = 8 if Lm >= 65536 else 1
    k_total = k_auto if k_auto % world == 0 else world
    variant = 2 if (F // T > 8192 and T <= 64) else 1
    lanes = 1 if Lm >= 49152 else (8 if Lm > 8192 else 32)         # belief_kernel: the lane count fixes the landmark summation order
    return dict(tile_edges=T, lmk_block=lblock, kernel_variant=variant, belief_lanes=lanes), k_total


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

    def __init__(self, sub: BALProblem, configs, device, stream, belief_lanes=0, **kw):
        import torch
        from .engine import BAEngine
        self.eng = BAEngine(sub.cam_id, sub.lmk_id, sub.z, sub.cam_means, sub.lmk_means, sub.K4, configs, device=device,
                            stream=stream, **kw)
        if belief_lanes:
            self.eng.tune(L.TUNE_BELIEF_LANES, belief_lanes)
        ptr, nbytes = self.eng.device_ptr(L.F_CAM_PARTIAL)
        self._partial = torch.as_tensor(_DevArray(ptr, nbytes // 8), device=f"cuda:{device}")
        self.partials_per_rank = self.eng.lmk_chunks          # chunk sums this rank contributes to the exchange
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
        self.eng.cam_update(gathered.data_ptr(), world * self.partials_per_rank)

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
                 engine_factory=None, torch_stream=None, **engine_kw):
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
        if engine_factory is None and world > 1:
            layout, k_total = global_layout(prob, world)
            k_local = k_total // world
            for k, v in layout.items():
                engine_kw.setdefault(k, v)
            engine_kw.setdefault("chunks", (k_local, rank * k_local, k_total, self.lmk_range[0], prob.n_points))
        factory = engine_factory or (lambda s, c: CudaEngineAdapter(s, c, device, stream, **engine_kw))
        self.adapter = factory(sub, configs)
        self._gather = self.adapter.new_gather_buffer(world) if world > 1 else None
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

    def get_means(self):
        """All belief means in variable order (keyframes, then landmarks) on every rank."""
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
        self.adapter.close()
