import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


def golden_configs(G):
    def conv(v):
        if v == "None":
            return None
        try:
            return float(v)
        except ValueError:
            return v
    cfg = {str(k): conv(str(v)) for k, v in zip(G["cfg_keys"], G["cfg_vals"])}
    for k in ("num_undamped_iters", "min_linear_iters"):
        cfg[k] = int(cfg[k])
    return cfg


def golden_problem(G):
    from gbp_b200 import balio
    return balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])


def relerr(a, b):
    """max |a - b| / max |b| over the WHOLE table (a table-level norm: small entries are measured against the largest one;
    use relerr_rows where every variable has to be right on its own scale)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def relerr_rows(a, b):
    """Per-variable relative error: max over rows v of ||a_v - b_v||_2 / ||b_v||_2 (Frobenius norm for matrix rows)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    return float(np.max(np.linalg.norm(a - b, axis=1) / den))


REF_COPY = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="session")
def built_library():
    from gbp_b200 import build
    return build.build_library()
