"""ctypes binding of libgbp_b200.so (include/gbp_b200.h).

There is NO CPU implementation behind this module: if the CUDA library is missing, or no
CUDA device is visible when a graph is created, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

ABI_VERSION = 3
COMM_ID_BYTES = 128

# gbp_field
(F_CAM_BELIEF, F_LMK_BELIEF, F_CAM_PRIOR, F_LMK_PRIOR, F_MSG_CAM, F_MSG_LMK, F_LINPOINT, F_ITERS,
 F_FLAGS, F_ADAPTIVE_VAR, F_MEASUREMENT, F_JACOBIAN_B, F_ADJ, F_FILE_INDEX, F_CAM_PARTIAL, F_CAM_MU, F_LMK_MU) = range(17)

# stages
ST_ROBUSTIFY, ST_RELIN, ST_MESSAGES, ST_BELIEFS, ST_LOCAL_DAMPING, ST_DEFER_LANDMARKS = 1, 2, 4, 8, 16, 32

LOSS_CODES = {None: 0, "huber": 1, "constant": 2}
TUNE_PREFETCH_TILES, TUNE_BELIEF_LANES = 3, 5

# field -> (index kind, dtype, row width)
FIELD_SHAPES = {
    F_CAM_BELIEF: ("C", np.float64, 33), F_LMK_BELIEF: ("L", np.float64, 12),
    F_CAM_PRIOR: ("C", np.float64, 27), F_LMK_PRIOR: ("L", np.float64, 9),
    F_MSG_CAM: ("F", np.float64, 27), F_MSG_LMK: ("F", np.float64, 9),
    F_LINPOINT: ("F", np.float64, 9), F_ITERS: ("F", np.int32, 1), F_FLAGS: ("F", np.int32, 1),
    F_ADAPTIVE_VAR: ("F", np.float64, 1), F_MEASUREMENT: ("F", np.float64, 2),
    F_JACOBIAN_B: ("F", np.float64, 20), F_ADJ: ("F", np.int32, 2), F_FILE_INDEX: ("F", np.int32, 1),
    F_CAM_PARTIAL: ("C", np.float64, 27), F_CAM_MU: ("C", np.float64, 6), F_LMK_MU: ("L", np.float64, 3),
}


class GbpError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from gbp_last_error)."""

    def __init__(self, status, message):
        super().__init__(f"libgbp_b200 status {status}: {message}")
        self.status = status


class GbpConfig(C.Structure):
    _fields_ = [("gauss_noise_std", C.c_double), ("eta_damping", C.c_double), ("beta", C.c_double),
                ("Nstds", C.c_double), ("num_undamped_iters", C.c_int32), ("min_linear_iters", C.c_int32),
                ("loss", C.c_int32), ("tile_edges", C.c_int32), ("lmk_block", C.c_int32), ("kernel_variant", C.c_int32),
                ("lmk_chunks", C.c_int32), ("lmk_chunk_first", C.c_int32), ("lmk_chunks_total", C.c_int32), ("reserved0", C.c_int32),
                ("lmk_first", C.c_int64), ("lmk_total", C.c_int64)]


EXPORTS = [
    "gbp_last_error", "gbp_abi_version", "gbp_device_count", "gbp_ba_create", "gbp_ba_destroy", "gbp_ba_reset", "gbp_ba_sizes", "gbp_ba_layout",
    "gbp_ba_prior_scan", "gbp_ba_generate_priors", "gbp_ba_set_priors", "gbp_ba_scale_priors",
    "gbp_ba_sweep_local", "gbp_ba_landmark_update", "gbp_ba_cam_update", "gbp_ba_iterate", "gbp_ba_update_beliefs", "gbp_ba_metrics",
    "gbp_ba_snapshot_layout", "gbp_ba_snapshot_async", "gbp_ba_snapshot_wait", "gbp_ba_iterate_snapshot", "gbp_host_alloc", "gbp_host_free", "gbp_ba_read", "gbp_ba_write", "gbp_ba_fill_iters", "gbp_ba_device_ptr", "gbp_ba_set_params",
    "gbp_ba_synchronize", "gbp_ba_time_iterations", "gbp_ba_launch_count", "gbp_reprojection_eval", "gbp_plan_create", "gbp_plan_sizes", "gbp_plan_copy", "gbp_plan_chunks", "gbp_plan_destroy", "gbp_bal_open", "gbp_bal_sizes", "gbp_bal_copy", "gbp_bal_close",
    "gbp_cache_configure", "gbp_cache_stats", "gbp_ba_tune",
    "gbp_lin_create", "gbp_lin_destroy", "gbp_lin_set_messages", "gbp_lin_update_beliefs", "gbp_lin_iterate", "gbp_lin_energy", "gbp_lin_read",
    "gbp_lin_joint_solve", "gbp_lin_launch_count",
    "gbp_comm_unique_id", "gbp_comm_version", "gbp_comm_create", "gbp_comm_destroy", "gbp_ba_attach_comm", "gbp_ba_detach_comm", "gbp_ba_comm_info", "gbp_ba_exchange",
]

_lib = None


def load():
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found.  gbp_b200 has no CPU fallback: build the CUDA library with "
            f"`python -m gbp_b200.build` (needs nvcc) before using the BA engine.")
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.gbp_last_error.restype = C.c_char_p
    lib.gbp_abi_version.restype = C.c_int
    lib.gbp_device_count.restype = C.c_int
    lib.gbp_ba_create.argtypes = [C.POINTER(GbpConfig), C.c_int32, C.c_int32, C.c_int64, vp, vp, vp, vp, vp, vp,
                                  C.c_int, vp, C.POINTER(vp)]
    lib.gbp_ba_destroy.argtypes = [vp]
    lib.gbp_ba_reset.argtypes = [vp]
    lib.gbp_ba_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.gbp_ba_layout.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.gbp_ba_prior_scan.argtypes = [vp, vp]
    lib.gbp_ba_generate_priors.argtypes = [vp, C.c_double, vp]
    lib.gbp_ba_set_priors.argtypes = [vp, vp, vp]
    lib.gbp_ba_scale_priors.argtypes = [vp, C.c_double]
    lib.gbp_ba_sweep_local.argtypes = [vp, C.c_int]
    lib.gbp_ba_landmark_update.argtypes = [vp]
    lib.gbp_ba_cam_update.argtypes = [vp, vp, C.c_int]
    lib.gbp_ba_iterate.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.gbp_ba_update_beliefs.argtypes = [vp]
    lib.gbp_ba_metrics.argtypes = [vp, dp]
    lib.gbp_ba_snapshot_layout.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.gbp_ba_snapshot_async.argtypes = [vp, vp]
    lib.gbp_ba_snapshot_wait.argtypes = [vp]
    lib.gbp_ba_iterate_snapshot.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.gbp_host_alloc.argtypes = [C.c_size_t]
    lib.gbp_host_alloc.restype = vp
    lib.gbp_host_free.argtypes = [vp]
    lib.gbp_host_free.restype = None
    lib.gbp_ba_read.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.gbp_ba_write.argtypes = [vp, C.c_int, vp, C.c_size_t]
    lib.gbp_ba_fill_iters.argtypes = [vp, C.c_int32]
    lib.gbp_ba_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.gbp_ba_set_params.argtypes = [vp, C.c_double, C.c_double, C.c_int32, C.c_int32]
    lib.gbp_ba_synchronize.argtypes = [vp]
    lib.gbp_ba_time_iterations.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                           C.POINTER(C.c_float)]
    lib.gbp_ba_launch_count.argtypes = [vp]
    lib.gbp_ba_launch_count.restype = C.c_int64
    lib.gbp_reprojection_eval.argtypes = [vp, C.c_int64, vp, C.c_int, vp, vp]
    lib.gbp_plan_create.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int64, vp, vp, C.POINTER(vp)]
    lib.gbp_plan_chunks.argtypes = [vp, vp, vp]
    lib.gbp_plan_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.gbp_plan_copy.argtypes = [vp] * 10
    lib.gbp_plan_destroy.argtypes = [vp]
    lib.gbp_plan_destroy.restype = None
    lib.gbp_bal_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.gbp_bal_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.gbp_bal_copy.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.gbp_bal_close.argtypes = [vp]
    lib.gbp_bal_close.restype = None
    lib.gbp_cache_configure.argtypes = [C.c_int32, C.c_int64]
    lib.gbp_cache_stats.argtypes = [C.POINTER(C.c_int64)]
    lib.gbp_ba_tune.argtypes = [vp, C.c_int, C.c_int64]
    lib.gbp_lin_create.argtypes = [C.c_int32, C.c_int32, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_int, vp, C.POINTER(vp)]
    lib.gbp_lin_destroy.argtypes = [vp]
    lib.gbp_lin_set_messages.argtypes = [vp, vp, vp]
    lib.gbp_lin_update_beliefs.argtypes = [vp]
    lib.gbp_lin_iterate.argtypes = [vp, C.c_int]
    lib.gbp_lin_energy.argtypes = [vp, dp]
    lib.gbp_lin_read.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.gbp_lin_joint_solve.argtypes = [vp, vp, vp, vp, vp]
    lib.gbp_lin_launch_count.argtypes = [vp]
    lib.gbp_lin_launch_count.restype = C.c_int64
    lib.gbp_comm_unique_id.argtypes = [vp]
    lib.gbp_comm_version.restype = C.c_int
    lib.gbp_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.gbp_comm_destroy.argtypes = [vp]
    lib.gbp_ba_attach_comm.argtypes = [vp, vp]
    lib.gbp_ba_detach_comm.argtypes = [vp]
    lib.gbp_ba_comm_info.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.gbp_ba_exchange.argtypes = [vp]
    if lib.gbp_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libgbp_b200 ABI {lib.gbp_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def prefer_bundled_nccl():
    """Point the library's dlopen at the NCCL that PyTorch bundles (nvidia/nccl/lib/libnccl.so.2) when there is one: a process
    cannot hold two different libnccl.so.2, and torch needs its own.  No-op when $GBP_NCCL_LIB is set or there is no bundled copy."""
    if os.environ.get("GBP_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["GBP_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def comm_version():
    """NCCL version code the library would use (0: no libnccl)."""
    prefer_bundled_nccl()
    return int(load().gbp_comm_version())


def check(status):
    if status != 0:
        raise GbpError(status, load().gbp_last_error().decode("utf-8", "replace"))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# Page-locked blocks are recycled: cudaHostAlloc / cudaFreeHost cost milliseconds and synchronise the device.
_PINNED_POOL = {}          # nbytes -> [ptr, ...]
_PINNED_POOL_CAP = 256 << 20
_pinned_pool_bytes = 0


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; returns it to the pool when the last array using it dies."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes
        self.buf = (C.c_char * nbytes).from_address(ptr)

    def __del__(self):
        global _pinned_pool_bytes
        try:
            if not self.ptr or _lib is None:
                return
            if _pinned_pool_bytes + self.nbytes <= _PINNED_POOL_CAP:
                _PINNED_POOL.setdefault(self.nbytes, []).append(self.ptr)
                _pinned_pool_bytes += self.nbytes
            else:
                _lib.gbp_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (fast asynchronous device->host copies); ordinary memory
    when page-locked memory is unavailable."""
    global _pinned_pool_bytes
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n == 0:
        return np.empty(shape, dtype=dtype)
    n_alloc = (n + 4095) & ~4095
    free = _PINNED_POOL.get(n_alloc)
    if free:
        p = free.pop()
        _pinned_pool_bytes -= n_alloc
    else:
        p = load().gbp_host_alloc(n_alloc)
    if not p:
        return np.empty(shape, dtype=dtype)
    blk = _PinnedBlock(p, n_alloc)
    arr = np.frombuffer(blk.buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return _attach(arr, blk)


class _OwnedArray(np.ndarray):
    """ndarray subclass that only adds a slot to keep the pinned block alive."""
    _gbp_block = None

    def __array_finalize__(self, obj):
        self._gbp_block = getattr(obj, "_gbp_block", None)


def _attach(arr, blk):
    out = arr.view(_OwnedArray)
    out._gbp_block = blk
    return out
