"""Multi-GPU GBP sweep: the factor graph is cut by LANDMARK across the ranks of one box.

Every factor lives with its landmark, so all landmarks are interior (their incoming messages
are local) and the boundary variables are the keyframes, replicated on every rank.  Per
synchronous iteration each rank runs the sweep over its own edges, reduces its local
factor->keyframe messages to (eta, Lambda) sums per keyframe and landmark chunk, and the
ranks exchange those sums with ONE all-gather; every rank then adds them in chunk order and the prior, so the
keyframe beliefs are bit-identical everywhere.  ARE / energy need a 3-scalar all-reduce only when the client asks.

The exchange itself is NOT in this file: the CUDA library calls NCCL (`gbp_ba_attach_comm`, gbp_b200/csrc/gbp_dist.cu.inc;
one process per GPU, the all-gather on a high-priority side stream overlapping the landmark belief update, the whole
iteration ONE CUDA-graph replay per rank), so a C or C++ client shards exactly like this Python one.  What this module
does: cut the problem (`local_problem`), choose the layout from the GLOBAL sizes (`global_layout`), carry the NCCL unique
id from rank 0 to the other ranks over the caller's `torch.distributed` group (any backend), and concatenate the means.

Bit-identical across the NUMBER of GPUs as well: the engine forms the keyframe-side sums per landmark CHUNK (gbp_config:
up to 8 chunks of at least 125 000 consecutive landmarks) and adds the chunk sums in chunk order; a rank holds whole chunks of
that global chunking and lays out exactly the tiles the single-GPU plan has for them (tile size, landmark blocks and kernel
build are chosen from the GLOBAL sizes).  So 1, 2, 4 and 8 GPUs run the same floating-point operations in the same order
(whenever the automatic chunk count is a multiple of the number of GPUs -- 8 chunks from 1 M landmarks on; otherwise the run uses
one chunk per rank and equals the single-GPU engine created with that chunking, `chunks=(world, 0, world, 0, n_landmarks)`).

(A collective-free exchange -- every rank storing its partial sums into the other ranks' buffers over NVLink through CUDA IPC
mappings, per-CTA flags -- was built and run on 2 and 8 GPUs in round 2: same bits, same speed as the all-gather within 1 %
(0.2143 vs 0.2130 ms per iteration at 8 GPUs), so it was removed again; profiles/r2e_bench_n8_*.json.)

The reference has no distributed code; this is new functionality behind the same
`synchronous_iteration` surface.

The compute engine is injectable (`engine_factory`): an engine WITHOUT its own exchange (`native_exchange` False, e.g. the
oracle-backed stand-in of tests/test_dist_cpu.py) is driven through the host-side schedule below (local sweep ->
`dist.all_gather_into_tensor` -> keyframe update), which is how the partition / merge logic is tested on CPU with gloo and
world_size 2.  The product always uses the CUDA engine and its native exchange.
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
    k_auto = 1
    while k_auto < 8 and Lm // (2 * k_auto) >= 125000:      # auto_chunks() of gbp_ba.cu
        k_auto *= 2
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


_COMMS = {}     # (device, rank, world) -> engine.Communicator of this process


def communicator(dist, rank, world, device):
    """This process's NCCL communicator for the library (created once: ncclCommInitRank costs seconds; graphs come and go).
    Collective on first use: rank 0 makes the unique id, `dist.broadcast_object_list` (any backend) carries its 128 bytes."""
    from .engine import Communicator
    key = (int(device), int(rank), int(world))
    c = _COMMS.get(key)
    if c is None:
        box = [Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        c = _COMMS[key] = Communicator(box[0], rank, world, device)
    return c


def shutdown():
    """Destroy the cached communicators (every graph attached to them must be closed); call before the process group goes away."""
    for key in list(_COMMS):
        _COMMS.pop(key).destroy()


class CudaEngineAdapter:
    """One rank's CUDA engine.  With world > 1 the library's own NCCL communicator is attached: iterate / update_beliefs /
    generate_priors / metrics are then collective calls that already cover the whole graph (`native_exchange`)."""

    native_exchange = True

    def __init__(self, sub: BALProblem, configs, device, stream, rank=0, world=1, dist=None, belief_lanes=0, **kw):
        from .engine import BAEngine
        self.eng = BAEngine(sub.cam_id, sub.lmk_id, sub.z, sub.cam_means, sub.lmk_means, sub.K4, configs, device=device,
                            stream=stream, **kw)
        if belief_lanes:
            self.eng.tune(L.TUNE_BELIEF_LANES, belief_lanes)
        if world > 1:
            self.eng.attach_comm(communicator(dist, rank, world, device))

    C = property(lambda self: self.eng.C)
    L = property(lambda self: self.eng.L)
    F = property(lambda self: self.eng.F)

    def generate_priors(self, weaker):
        self.eng.generate_priors(weaker)                     # keyframe maxima over all ranks inside the library

    def scale_priors(self, f):
        self.eng.scale_priors(f)

    def iterate(self, n, robustify, local_relin):
        self.eng.iterate(n, robustify=robustify, local_relin=local_relin)

    def update_beliefs(self):
        self.eng.update_beliefs()

    def metrics(self):
        a, e, n = self.eng.metrics()                         # sums over the whole graph when a communicator is attached
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
        """stream: raw cudaStream_t of the engine (None = its own); torch_stream: the same given as a torch.cuda.Stream."""
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs an initialised torch.distributed module")
        if torch_stream is not None:
            if stream is not None and int(stream) != int(torch_stream.cuda_stream):
                raise ValueError("stream and torch_stream name different CUDA streams")
            stream = torch_stream.cuda_stream
        self.rank, self.world, self.dist = rank, world, dist
        self.n_iterations = 0        # synchronous iterations applied to the state since creation / reset
        self.F_total, self.L_total, self.C = prob.n_edges, prob.n_points, prob.n_keyframes
        sub, self.local_measurements, self.lmk_range = local_problem(prob, rank, world)
        if engine_factory is None:
            if world > 1:
                layout, k_total = global_layout(prob, world)
                k_local = k_total // world
                for k, v in layout.items():
                    engine_kw.setdefault(k, v)
                engine_kw.setdefault("chunks", (k_local, rank * k_local, k_total, self.lmk_range[0], prob.n_points))
            self.adapter = CudaEngineAdapter(sub, configs, device, stream, rank=rank, world=world, dist=dist, **engine_kw)
        else:
            self.adapter = engine_factory(sub, configs)
        self.native = bool(getattr(self.adapter, "native_exchange", False))
        self._gather = self.adapter.new_gather_buffer(world) if (world > 1 and not self.native) else None

    @property
    def engine(self):
        return self.adapter.eng

    # ------------------------------------------------------------------ host-side exchange (engines without their own)
    def _exchange_and_update(self):
        """keyframe partial sums -> all ranks (one all-gather); the landmark beliefs need no communication; then the
        partials in rank order + prior."""
        a = self.adapter
        work = self.dist.all_gather_into_tensor(self._gather, a.partial_tensor(), async_op=True)
        a.landmark_update()
        work.wait()
        a.apply_gathered(self._gather, self.world)

    # ------------------------------------------------------------------ API
    def generate_priors_var(self, weaker_factor=100):
        """gbp/gbp_ba.py:20-34 with the per-keyframe maximum taken over all ranks."""
        a = self.adapter
        if self.native:
            a.generate_priors(weaker_factor)
            return
        cam_max = a.prior_scan()
        if self.world > 1:
            self.dist.all_reduce(cam_max, op=self.dist.ReduceOp.MAX)
        a.generate_priors(weaker_factor, cam_max)

    def weaken_priors(self, f):
        self.adapter.scale_priors(f)

    def update_all_beliefs(self):
        if self.native:
            self.adapter.update_beliefs()
        else:
            self.adapter.sweep_local(L.ST_BELIEFS | L.ST_DEFER_LANDMARKS)
            self._exchange_and_update()

    def synchronous_iteration(self, local_relin=True, robustify=False):
        """gbp/gbp.py:86-92 over the partitioned graph: local sweep -> one all-gather -> keyframe beliefs."""
        self.n_iterations += 1
        if self.native:
            self.adapter.iterate(1, robustify, local_relin)
            return
        st = L.ST_MESSAGES | L.ST_BELIEFS | L.ST_DEFER_LANDMARKS
        if robustify:
            st |= L.ST_ROBUSTIFY
        if local_relin:
            st |= L.ST_RELIN | L.ST_LOCAL_DAMPING
        self.adapter.sweep_local(st)
        self._exchange_and_update()

    def iterate(self, n, local_relin=True, robustify=False):
        """n synchronous iterations; on the CUDA engine ONE call into the library (graph replays, no Python in between)."""
        if self.native:
            self.adapter.iterate(n, robustify, local_relin)
            self.n_iterations += n
        else:
            for _ in range(n):
                self.synchronous_iteration(local_relin=local_relin, robustify=robustify)

    def capture(self, local_relin=True, robustify=False):
        """Kept for callers of the round-1 interface: the library replays every iteration from its own CUDA graph (sweep, chunk
        sums, all-gather, both belief updates), so there is nothing to capture here.  True when iterations are graph replays."""
        return self.native

    def fill_iters(self, value):
        self.adapter.fill_iters(value)

    def reset(self):
        """Back to the state right after construction (gbp_ba_reset on every rank): zero messages and priors."""
        self.adapter.reset()
        self.n_iterations = 0

    def metrics(self):
        """(ARE, energy, number of factors with iters_since_relin == 0) over the WHOLE graph."""
        m = self.adapter.metrics()
        if self.world > 1 and not self.native:
            import torch
            t = torch.from_numpy(np.asarray(m, dtype=np.float64).copy())
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            m = t.numpy()
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
            self.dist.all_gather_object(parts, np.asarray(lmk))
            lmk = np.concatenate(parts, axis=0)
        return np.concatenate([cam, np.asarray(lmk).ravel()])

    def close(self):
        self.adapter.close()
