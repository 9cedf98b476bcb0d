"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol
declared in include/gbp_b200.h; it refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gbp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gbp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    from gbp_b200 import _lib
    lib = ctypes.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gbp_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.load().gbp_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback(built_library):
    """Without a device, creating a graph must fail loudly (status GBP_ERR_NO_DEVICE)."""
    from gbp_b200 import _lib
    from gbp_b200.engine import BAEngine
    lib = _lib.load()
    if lib.gbp_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    cfg = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
    with pytest.raises(_lib.GbpError) as ei:
        BAEngine([0], [0], [[1.0, 2.0]], np.zeros((1, 6)), np.zeros((1, 3)), [500, 500, 320, 240], cfg)
    assert ei.value.status == 3 and "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_reported(built_library):
    from gbp_b200 import _lib
    lib = _lib.load()
    assert lib.gbp_ba_sizes(None, None) != 0
    assert b"null" in lib.gbp_last_error()
    assert lib.gbp_ba_destroy(None) == 0


def test_product_does_not_import_oracle():
    """The product package must never import, include or execute anything under oracle/."""
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s*[\"<].*oracle)|oracle[./]gbp_oracle|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gbp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references the oracle"
