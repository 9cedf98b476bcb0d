"""ctypes wrapper of oracle/gbp_oracle.c (plain C + OpenMP restatement of the reference sweep).

TEST INFRASTRUCTURE ONLY: used by tests/ and by the CPU-baseline legs of bench.py.  `build()` compiles it with
gcc into oracle/_build/ (git-ignored)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gbp_oracle.c")
LIB = os.path.join(HERE, "_build", "libgbp_oracle.so")
LOSS = {None: 0, "huber": 1, "constant": 2}
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fopenmp", "-std=c11", "-shared", "-fPIC", SRC, "-lm", "-o", LIB], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        vp = C.c_void_p
        lib.gbpo_create.restype = vp
        lib.gbpo_create.argtypes = [C.c_int, C.c_int, C.c_long, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_int, C.c_int, C.c_int]
        lib.gbpo_destroy.argtypes = [vp]
        lib.gbpo_generate_priors.argtypes = [vp, C.c_double]
        lib.gbpo_weaken_priors.argtypes = [vp, C.c_double]
        lib.gbpo_update_beliefs.argtypes = [vp]
        lib.gbpo_iterate.argtypes = [vp, C.c_int, C.c_int]
        lib.gbpo_metrics.argtypes = [vp, C.POINTER(C.c_double)]
        lib.gbpo_fill_iters.argtypes = [vp, C.c_int]
        lib.gbpo_field.restype = C.POINTER(C.c_double)
        lib.gbpo_field.argtypes = [vp, C.c_int]
        lib.gbpo_iters.restype = C.POINTER(C.c_int)
        lib.gbpo_iters.argtypes = [vp]
        lib.gbpo_threads.restype = C.c_int
        lib.gbpo_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def set_threads(n):
    """Number of OpenMP threads for subsequent calls (torchrun exports OMP_NUM_THREADS=1)."""
    load().gbpo_set_threads(int(n))


class COracle:
    """Same surface as oracle.gbp_oracle.BAOracle for the calls the tests and bench.py make."""

    FIELDS = {"cam_mu": (0, "C", 6), "lmk_mu": (1, "L", 3), "cam_eta": (2, "C", 6), "lmk_eta": (3, "L", 3),
              "cam_lam": (4, "C", 36), "lmk_lam": (5, "L", 9), "msg_cam_eta": (6, "F", 6), "msg_cam_lam": (7, "F", 36),
              "msg_lmk_eta": (8, "F", 3), "msg_lmk_lam": (9, "F", 9), "linpoint": (10, "F", 9), "adaptive_var": (11, "F", 1),
              "factor_damping": (12, "F", 1), "factor_eta": (13, "F", 9), "factor_lam": (14, "F", 81),
              "cam_prior_lam": (15, "C", 36), "lmk_prior_lam": (16, "L", 9)}

    def __init__(self, cam_id, lmk_id, z, cam0, lmk0, K4, configs):
        self._lib = load()
        a = [np.ascontiguousarray(cam_id, np.int32), np.ascontiguousarray(lmk_id, np.int32),
             np.ascontiguousarray(z, np.float64), np.ascontiguousarray(cam0, np.float64),
             np.ascontiguousarray(lmk0, np.float64), np.ascontiguousarray(K4, np.float64)]
        self.C, self.L, self.F = len(a[3]), len(a[4]), len(a[0])
        c = configs
        self._h = self._lib.gbpo_create(self.C, self.L, self.F, *[x.ctypes.data_as(C.c_void_p) for x in a],
                                        float(c["gauss_noise_std"]), float(c["eta_damping"]), float(c["beta"]),
                                        float(c.get("Nstds", 3.0)), int(c["num_undamped_iters"]), int(c["min_linear_iters"]),
                                        LOSS[c.get("loss", None)])
        self.threads = int(self._lib.gbpo_threads())

    def __del__(self):
        try:
            self._lib.gbpo_destroy(self._h)
        except Exception:
            pass

    def generate_priors_var(self, weaker_factor=100.0):
        self._lib.gbpo_generate_priors(self._h, float(weaker_factor))

    def weaken_priors(self, f):
        self._lib.gbpo_weaken_priors(self._h, float(f))

    def update_all_beliefs(self):
        self._lib.gbpo_update_beliefs(self._h)

    def synchronous_iteration(self, local_relin=True, robustify=False):
        self._lib.gbpo_iterate(self._h, int(bool(robustify)), int(bool(local_relin)))

    def metrics(self):
        out = (C.c_double * 3)()
        self._lib.gbpo_metrics(self._h, out)
        return float(out[0]), float(out[1]), int(round(out[2]))

    def are(self):
        return self.metrics()[0]

    def energy(self):
        return self.metrics()[1]

    def n_relinearising(self):
        return self.metrics()[2]

    def fill_iters(self, v):
        self._lib.gbpo_fill_iters(self._h, int(v))

    @property
    def iters_since_relin(self):
        return np.ctypeslib.as_array(self._lib.gbpo_iters(self._h), shape=(self.F,)).copy()

    def __getattr__(self, name):
        if name in COracle.FIELDS:
            which, kind, w = COracle.FIELDS[name]
            n = {"C": self.C, "L": self.L, "F": self.F}[kind]
            a = np.ctypeslib.as_array(self._lib.gbpo_field(self._h, which), shape=(n * w,)).copy()
            if w in (36, 9, 81) and name.endswith("lam"):
                d = int(round(w ** 0.5))
                return a.reshape(n, d, d)
            return a.reshape(n, w) if w > 1 else a
        raise AttributeError(name)


def run_ba_loop(o: COracle, n_iters, weaker_factor=50.0, float_impl=False, on_iter=None, final_weaker=100.0, n_weak=5):
    """The body of ba.py:75-105 (without the viewer)."""
    o.generate_priors_var(weaker_factor)
    o.update_all_beliefs()
    wf = np.log10(final_weaker) / n_weak
    are, en, nrel = [], [], []
    for i in range(n_iters):
        if float_impl and (i + 1) % 2 == 0 and i < n_weak * 2:
            o.weaken_priors(wf)
        if i == 3 or i == 8:
            o.fill_iters(1)
        a, e, n = o.metrics()
        are.append(a); en.append(e); nrel.append(n)
        o.synchronous_iteration(robustify=True, local_relin=True)
        if on_iter is not None:
            on_iter(i, o)
    a, e, n = o.metrics()
    are.append(a); en.append(e); nrel.append(n)
    return np.array(are), np.array(en), np.array(nrel)
