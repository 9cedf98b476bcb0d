"""BAEngine: thin object wrapper over one libgbp_b200 handle (one GPU, one stream).

All numerics run in the CUDA library; this class only marshals NumPy arrays across the C
ABI and unpacks the packed row formats.  It is what ``gbp_b200.ba.BAFactorGraph`` (the
mirror of the reference's ``gbp/gbp_ba.py``) drives.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_IU6 = np.triu_indices(6)
_IU3 = np.triu_indices(3)


def unpack_sym(packed, n):
    """[..., n(n+1)/2] packed upper triangle -> [..., n, n] symmetric."""
    iu = _IU6 if n == 6 else _IU3 if n == 3 else np.triu_indices(n)
    out = np.zeros(packed.shape[:-1] + (n, n))
    out[..., iu[0], iu[1]] = packed
    out[..., iu[1], iu[0]] = packed
    return out


def pack_sym(full):
    n = full.shape[-1]
    iu = _IU6 if n == 6 else _IU3 if n == 3 else np.triu_indices(n)
    return np.ascontiguousarray(full[..., iu[0], iu[1]])


class Communicator:
    """One NCCL communicator of this process (one GPU), created by the library (`gbp_comm_create`).  One per process and
    device is enough: it outlives the graphs that attach to it (creating one costs seconds)."""

    def __init__(self, unique_id, rank, nranks, device=0):
        if len(unique_id) != L.COMM_ID_BYTES:
            raise ValueError("unique_id must be the %d bytes of Communicator.unique_id()" % L.COMM_ID_BYTES)
        L.prefer_bundled_nccl()
        self._lib = L.load()
        self.handle = C.c_void_p()
        self.rank, self.nranks, self.device = int(rank), int(nranks), int(device)
        L.check(self._lib.gbp_comm_create(C.c_char_p(unique_id), int(rank), int(nranks), int(device), C.byref(self.handle)))

    @staticmethod
    def unique_id():
        """A fresh NCCL unique id (bytes); one rank makes it, every rank passes it to the constructor."""
        L.prefer_bundled_nccl()
        buf = C.create_string_buffer(L.COMM_ID_BYTES)
        L.check(L.load().gbp_comm_unique_id(buf))
        return buf.raw

    def destroy(self):
        if self.handle:
            L.check(self._lib.gbp_comm_destroy(self.handle))
            self.handle = C.c_void_p()


class BAEngine:
    def __init__(self, cam_id, lmk_id, z, cam_mu0, lmk_mu0, K4, configs, device=0, stream=None,
                 tile_edges=0, lmk_block=0, kernel_variant=0, chunks=None):
        """chunks: None = automatic chunking of this graph, or (lmk_chunks, lmk_chunk_first, lmk_chunks_total, lmk_first, lmk_total):
        this graph's share of a global landmark chunking (multi-GPU, see gbp_config in include/gbp_b200.h)."""
        lib = L.load()
        self._lib = lib
        self._h = None
        self._comm = None
        cam_id = np.ascontiguousarray(cam_id, dtype=np.int32)
        lmk_id = np.ascontiguousarray(lmk_id, dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        cam_mu0 = np.ascontiguousarray(cam_mu0, dtype=np.float64).reshape(-1, 6)
        lmk_mu0 = np.ascontiguousarray(lmk_mu0, dtype=np.float64).reshape(-1, 3)
        K4 = np.ascontiguousarray(K4, dtype=np.float64).reshape(4)
        if not (len(cam_id) == len(lmk_id) == len(z)):
            raise ValueError("cam_id, lmk_id and z must have one row per measurement")
        loss = configs.get("loss", None)
        if loss not in L.LOSS_CODES:
            raise ValueError(f"unknown loss {loss!r} (None, 'huber' or 'constant')")
        cfg = L.GbpConfig(float(configs["gauss_noise_std"]), float(configs["eta_damping"]), float(configs["beta"]),
                          float(configs.get("Nstds", 3.0)), int(configs["num_undamped_iters"]),
                          int(configs["min_linear_iters"]), L.LOSS_CODES[loss], int(tile_edges), int(lmk_block), int(kernel_variant),
                          *([0, 0, 0, 0, 0, 0] if chunks is None else [int(chunks[0]), int(chunks[1]), int(chunks[2]), 0, int(chunks[3]), int(chunks[4])]))
        self.cfg = cfg
        h = C.c_void_p()
        L.check(lib.gbp_ba_create(C.byref(cfg), len(cam_mu0), len(lmk_mu0), len(cam_id), L.ptr(cam_id), L.ptr(lmk_id),
                                  L.ptr(z), L.ptr(cam_mu0), L.ptr(lmk_mu0), L.ptr(K4), int(device),
                                  C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        sizes = (C.c_int64 * 6)()
        L.check(lib.gbp_ba_sizes(h, sizes))
        self.C, self.L, self.F, self.n_tiles, self.tile_edges, self.n_slots = [int(v) for v in sizes]
        lay = (C.c_int64 * 4)()
        L.check(lib.gbp_ba_layout(h, lay))
        # doubles per stored factor->keyframe message (27 full / 18 factored), L2 prefetch distance, sweep kernel build
        self.msg_cam_width, self.prefetch_tiles, self.sweep_variant = int(lay[0]), int(lay[1]), int(lay[2])
        self.lmk_chunks = int(lay[3])      # landmark chunks of the keyframe-side sums (GBP_F_CAM_PARTIAL holds lmk_chunks x C rows)
        self.K4 = K4
        self.device = int(device)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h is not None:
            self._lib.gbp_ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        L.check(self._lib.gbp_ba_reset(self._h))

    def tune(self, knob, value):
        """Engine tuning knobs (L.TUNE_*; measurement scripts): the L2 prefetch distance of the streaming build."""
        L.check(self._lib.gbp_ba_tune(self._h, int(knob), int(value)))
        if knob == L.TUNE_PREFETCH_TILES:
            self.prefetch_tiles = int(value) if self.sweep_variant == 2 else 0

    # ------------------------------------------------------------------ priors
    def prior_scan(self):
        out = np.zeros(self.C)
        L.check(self._lib.gbp_ba_prior_scan(self._h, L.ptr(out)))
        return out

    def generate_priors(self, weaker_factor, cam_max=None):
        cm = None if cam_max is None else np.ascontiguousarray(cam_max, dtype=np.float64)
        L.check(self._lib.gbp_ba_generate_priors(self._h, float(weaker_factor), L.ptr(cm)))

    def set_priors(self, cam_lam_packed, lmk_lam_packed):
        a = np.ascontiguousarray(cam_lam_packed, dtype=np.float64).reshape(self.C, 21)
        b = np.ascontiguousarray(lmk_lam_packed, dtype=np.float64).reshape(self.L, 6)
        L.check(self._lib.gbp_ba_set_priors(self._h, L.ptr(a), L.ptr(b)))

    def scale_priors(self, f):
        L.check(self._lib.gbp_ba_scale_priors(self._h, float(f)))

    # ------------------------------------------------------------------ sweeps
    def sweep_local(self, stages):
        L.check(self._lib.gbp_ba_sweep_local(self._h, int(stages)))

    def landmark_update(self):
        L.check(self._lib.gbp_ba_landmark_update(self._h))

    def cam_update(self, partials_dev_ptr=None, nranks=1):
        L.check(self._lib.gbp_ba_cam_update(self._h, C.c_void_p(partials_dev_ptr) if partials_dev_ptr else None,
                                            int(nranks)))

    # ------------------------------------------------------------------ multi-GPU (one process per GPU, NCCL called by the library)
    def attach_comm(self, comm):
        """Collective over the ranks of `comm` (a Communicator).  From now on iterate / update_beliefs exchange the keyframe chunk
        sums, generate_priors (without cam_max) and prior_scan take the keyframe maxima over all ranks and metrics() sums over the
        whole graph.  Detach or close this engine before the communicator is destroyed."""
        L.check(self._lib.gbp_ba_attach_comm(self._h, comm.handle))
        self._comm = comm          # keeps it alive

    def detach_comm(self):
        L.check(self._lib.gbp_ba_detach_comm(self._h))
        self._comm = None

    def comm_info(self):
        out = (C.c_int32 * 2)()
        L.check(self._lib.gbp_ba_comm_info(self._h, out))
        return int(out[0]), int(out[1])

    def exchange(self):
        L.check(self._lib.gbp_ba_exchange(self._h))

    def iterate(self, n_iters=1, robustify=False, local_relin=True):
        L.check(self._lib.gbp_ba_iterate(self._h, int(n_iters), int(bool(robustify)), int(bool(local_relin))))

    def update_beliefs(self):
        L.check(self._lib.gbp_ba_update_beliefs(self._h))

    def metrics(self):
        """(sum of |r|, energy, number of factors with iters_since_relin == 0) over the local edges."""
        out = (C.c_double * 3)()
        L.check(self._lib.gbp_ba_metrics(self._h, out))
        return float(out[0]), float(out[1]), int(round(out[2]))

    def snapshot_layout(self):
        out = (C.c_uint64 * 4)()
        L.check(self._lib.gbp_ba_snapshot_layout(self._h, out))
        return [int(v) for v in out]

    def snapshot_async(self, region):
        L.check(self._lib.gbp_ba_snapshot_async(self._h, L.ptr(region)))

    def iterate_snapshot(self, robustify, local_relin, region):
        L.check(self._lib.gbp_ba_iterate_snapshot(self._h, int(bool(robustify)), int(bool(local_relin)), L.ptr(region)))

    def snapshot_wait(self):
        L.check(self._lib.gbp_ba_snapshot_wait(self._h))

    def fill_iters(self, value):
        L.check(self._lib.gbp_ba_fill_iters(self._h, int(value)))

    def set_params(self, eta_damping, beta, num_undamped_iters, min_linear_iters):
        L.check(self._lib.gbp_ba_set_params(self._h, float(eta_damping), float(beta), int(num_undamped_iters),
                                            int(min_linear_iters)))

    def synchronize(self):
        L.check(self._lib.gbp_ba_synchronize(self._h))

    def time_iterations(self, n_iters, robustify=True, local_relin=True, per_kernel=False):
        a, b = C.c_float(), C.c_float()
        L.check(self._lib.gbp_ba_time_iterations(self._h, int(n_iters), int(bool(robustify)), int(bool(local_relin)),
                                                 int(bool(per_kernel)), C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def launch_count(self):
        return int(self._lib.gbp_ba_launch_count(self._h))

    def device_ptr(self, field):
        p, n = C.c_void_p(), C.c_size_t()
        L.check(self._lib.gbp_ba_device_ptr(self._h, int(field), C.byref(p), C.byref(n)))
        return p.value, n.value

    # ------------------------------------------------------------------ field access
    def _rows(self, kind):
        return {"C": self.C, "L": self.L, "F": self.F}[kind]

    def read(self, field, out=None):
        kind, dt, w = L.FIELD_SHAPES[field]
        n = self._rows(kind) * (self.lmk_chunks if field == L.F_CAM_PARTIAL else 1)      # chunk sums: chunks x C rows
        if out is None:
            out = L.pinned_empty((n, w), dt)
        assert out.dtype == dt and out.size == n * w and out.flags.c_contiguous
        L.check(self._lib.gbp_ba_read(self._h, int(field), L.ptr(out), out.nbytes))
        return out

    def write(self, field, arr):
        kind, dt, w = L.FIELD_SHAPES[field]
        n = self._rows(kind)
        a = np.ascontiguousarray(arr, dtype=dt)
        if a.size != n * w:
            raise ValueError(f"field {field}: expected {n}x{w} values, got {a.size}")
        L.check(self._lib.gbp_ba_write(self._h, int(field), L.ptr(a), a.nbytes))


def reprojection_eval(x, K4, device=0):
    """meas_fn / jac_fn of the reprojection factor evaluated by the CUDA kernels (parity tests)."""
    lib = L.load()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 9)
    K4 = np.ascontiguousarray(K4, dtype=np.float64).reshape(4)
    h = np.empty((len(x), 2))
    J = np.empty((len(x), 18))
    L.check(lib.gbp_reprojection_eval(L.ptr(x), len(x), L.ptr(K4), int(device), L.ptr(h), L.ptr(J)))
    return h, J.reshape(-1, 2, 9)


def compile_plan(cam_id, lmk_id, n_keyframes, n_landmarks, tile_edges=0, lmk_block=0, chunks=None):
    """The engine's storage order for a measurement list, from the native host graph compiler (`gbp_plan_*`; pure host
    code, no GPU needed): dict with T, tiles [(keyframe, count)], slot_of_factor, file_of_factor, adj, lmk_idx (per slot),
    lmk_ptr / lmk_slots (CSR by landmark over slots), cam_tile_ptr / cam_tiles (CSR by keyframe over tiles), n_chunks,
    tile_chunk (landmark chunk of every tile) and cam_chunk_ptr (per keyframe where its chunks start in cam_tiles)."""
    lib = L.load()
    cam_id = np.ascontiguousarray(cam_id, dtype=np.int32)
    lmk_id = np.ascontiguousarray(lmk_id, dtype=np.int32)
    if len(cam_id) != len(lmk_id):
        raise ValueError("cam_id and lmk_id must have one entry per measurement")
    p = C.c_void_p()
    if isinstance(chunks, int):
        chunks = (chunks, 0, chunks, 0, n_landmarks)
    ch = None if chunks is None else np.array(chunks, dtype=np.int64)
    L.check(lib.gbp_plan_create(int(tile_edges), int(lmk_block), L.ptr(ch), int(n_keyframes), int(n_landmarks), len(cam_id),
                                L.ptr(cam_id), L.ptr(lmk_id), C.byref(p)))
    try:
        sizes = (C.c_int64 * 6)()
        L.check(lib.gbp_plan_sizes(p, sizes))
        nc, nl, nf, n_tiles, T, n_slots = [int(v) for v in sizes]
        out = {"T": T, "n_tiles": n_tiles, "n_slots": n_slots,
               "tiles": np.zeros((n_tiles, 2), np.int32), "slot_of_factor": np.zeros(nf, np.int32),
               "file_of_factor": np.zeros(nf, np.int32), "adj": np.zeros((nf, 2), np.int32), "lmk_idx": np.zeros(n_slots, np.int32),
               "lmk_ptr": np.zeros(nl + 1, np.int32), "lmk_slots": np.zeros(nf, np.int32),
               "cam_tile_ptr": np.zeros(nc + 1, np.int32), "cam_tiles": np.zeros(n_tiles, np.int32)}
        L.check(lib.gbp_plan_copy(p, *[L.ptr(out[k]) for k in ("tiles", "slot_of_factor", "file_of_factor", "adj", "lmk_idx",
                                                              "lmk_ptr", "lmk_slots", "cam_tile_ptr", "cam_tiles")]))
        out["n_chunks"] = int(lib.gbp_plan_chunks(p, None, None))
        out["tile_chunk"] = np.zeros(n_tiles, np.int32)
        out["cam_chunk_ptr"] = np.zeros((nc, out["n_chunks"] + 1), np.int32)
        lib.gbp_plan_chunks(p, L.ptr(out["tile_chunk"]), L.ptr(out["cam_chunk_ptr"]))
        return out
    finally:
        lib.gbp_plan_destroy(p)
