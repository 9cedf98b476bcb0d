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
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1, 1000, 1_000_000)       # both arms print the same config object
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


def test_reference_arm_multi_gpu_line_matches_our_config():
    """N > 1: rank 0 times the C port on a bounded sample of the partitioned synthetic workload and prints OUR arm's config."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--synth-cams", "100", "--synth-lmks", "20000"], capture_output=True, text=True, check=True, cwd=ROOT, env=env,
                         timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong"
    assert d["config"] == bench.bench_config(2, 100, 20000) and "partitioned over 2 GPUs" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 1e5 and d["e2e"]["value"] == d["value"]


def test_roofline_entry_counts_both_byte_figures():
    """`achieved` / `frac` = the bytes the layout in use has to move / launch time (what the HBM delivers); SURVEY 8(d)'s
    layout-independent 696 B per factor is reported beside it as *_survey_bytes."""
    sys.path.insert(0, ROOT)
    import bench
    F, L, C = 10_000_000, 1_000_000, 1000
    _, survey = bench.b_alg(F, L, C)
    _, moved = bench.b_alg(F, L, C, 18)
    assert survey == 696 * F + 96 * L + 264 * C and moved == 552 * F + 96 * L + 264 * C
    r = bench.roofline_entry("w", 1.0, survey, moved, 6547.8, 5_599_662_000, 18)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6547.8
    assert abs(r["achieved"] - moved / 1e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / 6547.8) < 1e-12
    assert abs(r["achieved_survey_bytes"] - survey / 1e-3 / 1e9) < 1e-6 and r["frac"] < r["frac_survey_bytes"]
    assert r["algorithmic_bytes_per_launch"] == moved and r["survey_bytes_per_launch"] == survey
    assert r["traffic"] == 5_599_662_000 and r["ms_per_launch"] == 1.0 and "552" in r["bytes_note"]
    full = bench.roofline_entry("w", 1.2, survey, survey, 6547.8, None, 27)
    assert full["achieved"] == full["achieved_survey_bytes"] and full["traffic"] is None
    json.dumps(r)


def test_synthetic_section_builds_its_json_with_a_stub_engine(monkeypatch):
    """bench_synthetic_1gpu's bookkeeping (byte counts, roofline object, JSON-serialisable output) with the GPU engine, torch
    and the generator replaced by stubs: a typo there would cost the round its bench line."""
    import types
    sys.path.insert(0, ROOT)
    import bench
    import gbp_b200.dist as gdist
    import gbp_b200.synthetic as gsyn

    class Ev:
        def record(self): pass
        def elapsed_time(self, other): return 24.0

    class Eng:
        F, L, C, n_tiles, tile_edges = 10_000_000, 1_000_000, 1000, 158239, 64
        msg_cam_width, sweep_variant, prefetch_tiles = 18, 2, 600
        def launch_count(self): return 0
        def time_iterations(self, k, r, l, per_kernel=False): return 1.2 * k, 1.0 * k

    class PG:
        n_iterations = 223
        def __init__(self, *a, **kw): self.engine = Eng()
        def generate_priors_var(self, w): pass
        def update_all_beliefs(self): pass
        def synchronous_iteration(self, **kw): pass
        def metrics(self): return 2.3, 8.5e6, 0
        def close(self): pass

    monkeypatch.setattr(gdist, "PartitionedBAGraph", PG)
    monkeypatch.setattr(gsyn, "make_synthetic", lambda c, l, o, seed=0: types.SimpleNamespace(n_edges=10_000_000, n_points=1_000_000, n_keyframes=1000))
    args = types.SimpleNamespace(synth_cams=1000, synth_lmks=1_000_000, synth_iters=20, warmup=3, synth_sustained=200)
    ctx = types.SimpleNamespace(torch=types.SimpleNamespace(cuda=types.SimpleNamespace(synchronize=lambda: None)), local=0, stream=1,
                                hbm_peak=6547.8, events=lambda: (Ev(), Ev()))
    synth, roof = bench.bench_synthetic_1gpu(ctx, args)
    json.dumps({"synthetic": synth, "roofline": roof})
    assert abs(synth["ms_per_iteration"] - 1.2) < 1e-9 and synth["layout"]["msg_cam_doubles"] == 18
    assert abs(roof["ms_per_launch"] - 1.0) < 1e-12 and roof["frac_survey_bytes"] > 1.0 > roof["frac"] > 0.8
    assert roof["traffic"] == json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sweep_kernel/synthetic_1000_1000000_10000000"]["bytes"]


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
