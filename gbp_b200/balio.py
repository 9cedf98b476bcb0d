"""BAL-style problem files (format: reference data/README.md:5-16, parsed by
utils/read_balfile.py:4-37) -- reader, writer and a binary ``.npz`` container of the same
arrays for graphs too large for text.

    num_cameras num_landmarks num_measurements
    f_x f_y c_x c_y
    cam_index lmk_index x y            (num_measurements lines)
    one camera parameter per line      (6 per camera: T_cw translation, axis-angle rotation)
    one landmark coordinate per line   (3 per landmark)
"""
from __future__ import annotations

import os

import numpy as np


class BALProblem:
    """The 9-tuple returned by the reference's read_balfile, as arrays."""

    def __init__(self, cam_id, lmk_id, z, cam_means, lmk_means, K4):
        self.cam_id = np.ascontiguousarray(cam_id, dtype=np.int32)
        self.lmk_id = np.ascontiguousarray(lmk_id, dtype=np.int32)
        self.z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
        self.cam_means = np.ascontiguousarray(cam_means, dtype=np.float64).reshape(-1, 6)
        self.lmk_means = np.ascontiguousarray(lmk_means, dtype=np.float64).reshape(-1, 3)
        self.K4 = np.ascontiguousarray(K4, dtype=np.float64).reshape(4)

    @property
    def n_keyframes(self):
        return len(self.cam_means)

    @property
    def n_points(self):
        return len(self.lmk_means)

    @property
    def n_edges(self):
        return len(self.cam_id)

    @property
    def K(self):
        fx, fy, cx, cy = self.K4
        return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])

    def as_tuple(self):
        """Same order and types as utils/read_balfile.py:37."""
        return (self.n_keyframes, self.n_points, self.n_edges, self.cam_means, self.lmk_means, self.z,
                self.cam_id.tolist(), self.lmk_id.tolist(), self.K)


def read_bal(path) -> BALProblem:
    """BAL text file (native parser in libgbp_b200, `gbp_bal_open`) or the `.npz` container."""
    if str(path).endswith(".npz"):
        d = np.load(path)
        return BALProblem(d["cam_id"], d["lmk_id"], d["z"], d["cam_means"], d["lmk_means"], d["K4"])
    import ctypes as C
    from . import _lib as L
    if not os.path.exists(L.LIB_PATH):
        # reading a problem file is host-only work: a machine without the CUDA build (no nvcc) can still load and inspect
        # problems with the Python reader (same acceptance rules, ~8x slower); everything that computes still needs the library
        return read_bal_python(path)
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.gbp_bal_open(str(path).encode(), C.byref(h)))
    try:
        n = (C.c_int64 * 3)()
        L.check(lib.gbp_bal_sizes(h, n))
        nc, nl, nf = (int(v) for v in n)
        cam_id, lmk_id = np.empty(nf, np.int32), np.empty(nf, np.int32)
        z, cam, lmk, K4 = np.empty((nf, 2)), np.empty((nc, 6)), np.empty((nl, 3)), np.empty(4)
        L.check(lib.gbp_bal_copy(h, L.ptr(cam_id), L.ptr(lmk_id), L.ptr(z), L.ptr(cam), L.ptr(lmk), L.ptr(K4)))
    finally:
        lib.gbp_bal_close(h)
    return BALProblem(cam_id, lmk_id, z, cam, lmk, K4)


def read_bal_python(path) -> BALProblem:
    """Pure-Python reader with the same rules (cross-check of the native parser in the tests)."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    # header: skip blank lines and '# ...' comment lines (utils/read_balfile.py:7-11)
    i = 0
    while True:
        if i >= len(lines):
            raise ValueError(f"{path}: no header line found")
        tok = lines[i].split()
        if tok and tok[0] != "#":
            break
        i += 1
    n_cam, n_lmk, n_edges = (int(t) for t in lines[i].split())
    K4 = np.array([float(t) for t in lines[i + 1].split()[:4]])
    body = lines[i + 2:]
    need = n_edges + 6 * n_cam + 3 * n_lmk
    if len(body) < need:
        raise ValueError(f"{path}: expected {need} data lines, found {len(body)}")
    edge_tok = " ".join(body[:n_edges]).split()
    if len(edge_tok) == 4 * n_edges:
        e = np.array(edge_tok, dtype=np.float64).reshape(n_edges, 4)
    else:  # extra columns: keep the first four of every line like the reference does
        e = np.array([ln.split()[:4] for ln in body[:n_edges]], dtype=np.float64).reshape(n_edges, 4)
    par_lines = body[n_edges:need]
    par_tok = " ".join(par_lines).split()
    if len(par_tok) != len(par_lines):
        par_tok = [ln.split()[0] for ln in par_lines]
    par = np.array(par_tok, dtype=np.float64)
    return BALProblem(e[:, 0].astype(np.int32), e[:, 1].astype(np.int32), e[:, 2:4],
                      par[:6 * n_cam].reshape(n_cam, 6), par[6 * n_cam:].reshape(n_lmk, 3), K4)


def write_bal(path, prob: BALProblem, comments=()):
    """Text (round-trip exact: %.17g) or ``.npz`` depending on the extension."""
    if str(path).endswith(".npz"):
        np.savez(path, cam_id=prob.cam_id, lmk_id=prob.lmk_id, z=prob.z, cam_means=prob.cam_means,
                 lmk_means=prob.lmk_means, K4=prob.K4)
        return
    with open(path, "w") as f:
        for c in comments:
            f.write(f"# {c}\n")
        f.write("\n")
        f.write(f"{prob.n_keyframes} {prob.n_points} {prob.n_edges}\n")
        f.write(" ".join(f"{v:.17g}" for v in prob.K4) + "\n")
        for c, l, (x, y) in zip(prob.cam_id.tolist(), prob.lmk_id.tolist(), prob.z.tolist()):
            f.write(f"{c} {l}     {x:.17g} {y:.17g}\n")
        for v in prob.cam_means.ravel().tolist():
            f.write(f"{v:.17g}\n")
        for v in prob.lmk_means.ravel().tolist():
            f.write(f"{v:.17g}\n")
