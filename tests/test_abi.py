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


def test_cpp_shard_client_builds_against_the_header_and_fails_loudly_without_a_gpu(built_library, tmp_path):
    """examples/shard_client.cpp needs nothing but include/gbp_b200.h and the library (no NCCL header, no torch); on a host
    without a device it reports the library's error instead of computing anything on the CPU."""
    import subprocess
    from gbp_b200 import _lib, balio
    exe = str(tmp_path / "shard_client")
    lib_dir = os.path.dirname(built_library)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", os.path.join(ROOT, "examples", "shard_client.cpp"),
                    "-I" + os.path.join(ROOT, "include"), "-L" + lib_dir, "-lgbp_b200", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk_vsmall.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
    bal = str(tmp_path / "p.txt")
    balio.write_bal(bal, prob)
    res = subprocess.run([exe, "--rank", "0", "--nranks", "1", "--id-file", str(tmp_path / "id"), "--bal", bal, "--iters", "2"],
                         capture_output=True, text=True, timeout=120)
    if _lib.load().gbp_device_count() > 0:
        assert res.returncode == 0 and "after   2 ARE" in res.stdout, res.stderr
    else:
        assert res.returncode == 1 and "no CPU fallback" in res.stderr, (res.returncode, res.stderr)
    bad = subprocess.run([exe, "--rank", "3", "--nranks", "2", "--bal", bal], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 2 and "usage" in bad.stderr


def test_comm_api_without_a_device(built_library):
    """The multi-GPU entry points check their arguments before touching NCCL or a device."""
    from gbp_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.gbp_comm_create(None, 0, 1, 0, ctypes.byref(h)) == 1                          # GBP_ERR_INVALID: null id
    assert lib.gbp_comm_create(ctypes.c_char_p(b"\0" * 128), 2, 2, 0, ctypes.byref(h)) == 1   # rank out of range
    assert lib.gbp_comm_destroy(None) == 0
    assert lib.gbp_ba_attach_comm(None, None) != 0 and lib.gbp_ba_exchange(None) != 0
    assert _lib.comm_version() >= 20000                                                      # an NCCL 2.x is loadable in this image
    if lib.gbp_device_count() == 0:
        assert lib.gbp_comm_create(ctypes.c_char_p(b"\0" * 128), 0, 1, 0, ctypes.byref(h)) == 3   # GBP_ERR_NO_DEVICE


def test_header_is_plain_c(tmp_path):
    """include/gbp_b200.h is the boundary a C client binds to: it must compile as C99 on its own (no C++ in the signatures)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "gbp_b200.h"\nint main(void) { gbp_config c = {0}; (void)c; return GBP_COMM_ID_BYTES == 128 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), "-fsyntax-only", str(src)], check=True)
