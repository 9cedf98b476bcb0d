"""The driver-facing contracts that can be checked without a GPU: `bench.py --impl reference` prints ONE JSON line
with the agreed keys; `__graft_entry__.build()` compiles everything and every exported symbol resolves."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True, cwd=ROOT, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "gbp_messages_per_sec" and d["unit"] == "msgs/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert "fr1desk" in d["config"]["workload"] and d["config"]["msgs_per_step"] == 200 * 2 * 13298
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "msgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e5 and d["parity"]["max_rel_err_means_vs_reference_fixture"] < 1e-4


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_stdout_carries_only_the_json_line():
    """Libraries (NCCL's version banner) write to fd 1; bench.py keeps a private handle for its one JSON line."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.protect_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('python-level noise'); bench.emit('{\"ok\": 1}')" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert res.returncode == 0, res.stderr
    assert res.stdout == '{"ok": 1}\n' and "NCCL version" in res.stderr and "python-level noise" in res.stderr


def test_build_entry_point():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(os.path.join(ROOT, "gbp_b200", "lib", "libgbp_b200.so"))
    assert os.path.exists(os.path.join(ROOT, "oracle", "_build", "libgbp_oracle.so"))
    assert os.path.exists(os.path.join(ROOT, "tests", "host_harness", "_build", "libgbp_math_host.so"))


def test_sass_uses_the_bulk_copy_engine():
    """The shipped library really contains the TMA / mbarrier path (SASS mnemonics UBLKCP, SYNCS)."""
    import shutil
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gbp_b200", "lib", "libgbp_b200.so")], capture_output=True,
                          text=True, check=True).stdout
    assert "UBLKCP.S.G" in sass and "UBLKCP.G.S" in sass and "SYNCS.ARRIVE.TRANS64" in sass
    assert "UBLKPF.L2" in sass          # cp.async.bulk.prefetch.L2: the far-ahead prefetch of the large-graph build
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "gbp_b200", "lib", "libgbp_b200.so")],
                                       capture_output=True, text=True).stdout
