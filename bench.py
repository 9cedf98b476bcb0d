#!/usr/bin/env python
"""Benchmark of the GBP bundle-adjustment sweep (BASELINE.json metric: GBP messages / s per synchronous iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 -- headline workload = BASELINE configs[2], the configuration the metric is quoted on:
``ba.py --bal_file data/fr1desk.txt``, defaults, 200 synchronous iterations.  One STEP = one complete 200-iteration
solve from the initial state (including the client's `iters_since_relin = 1` resets at iterations 3 and 8,
ba.py:91-93); the graph state is reset and L2 is flushed between steps (both untimed), inside a step the 10 MB working
set is legitimately L2-resident, as in a real run.
  value  = GBP messages/s = 200 * 2F / device time of the step (CUDA events on the engine's stream)
  e2e    = the same step through the public Python API starting from pinned HOST arrays: create_ba_graph (graph compile
           + upload), priors, the 200 synchronous iterations with the two resets, final means read back to the host;
           wall clock with synchronisation; per-phase p50 / max in `e2e.phases_ms`.  `e2e_client_loop` is the same solve
           driven exactly like ba.py's loop body (are / energy / relinearisation count and the viewer's means read back
           to the host between every two sweeps); `unmodified_ba_py` is the wall time of the reference's own ba.py
           (staged copy under baseline/_ref) run against this engine in a subprocess.
The same run also measures the synthetic 1k-keyframe / 1M-landmark / 10M-factor graph (configs[3]; 5.8 GB streamed per
iteration, far larger than L2), which is where the HBM roofline is meaningful: `roofline` refers to the sweep kernel on
that graph, `roofline_fr1desk` to the (latency-bound, L2-resident) headline graph.

N > 1 -- fr1desk does not shard (10 MB); the path that shards is configs[4]: the synthetic graph cut by landmark over the
N GPUs with one exchange of keyframe partial sums per iteration.  The headline of an N-GPU line is therefore THAT graph
(`config.workload` says so), strong scaling: one step = 200 synchronous iterations from the initial state, value =
200 * 2F / device time (max over ranks).  Rank 0 also runs the single-GPU engine on the whole graph for the same 200
iterations in the same process (`parity_vs_1gpu`: agreement of the means and ARE, and the 1-GPU time on the same box, so
that the speed-up does not depend on comparing boxes).  The replicated fr1desk solve is reported under `fr1desk_replicas`.

``--impl reference`` times the CPU restatement of the reference algorithm (oracle/gbp_oracle.c: plain C + OpenMP on
every host thread) on the same workload; the UNMODIFIED NumPy reference itself (baseline/_ref, single-threaded Python)
is timed in the N = 1 run under `cpu_baseline.reference_numpy`.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_COPY = os.path.join(ROOT, "baseline", "_ref")

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
           eta_damping=0.4, prior_std_weaker_factor=50.0)
N_ITERS = 200
METRIC = "gbp_messages_per_sec"
UNIT = "msgs/s"
FR1 = dict(C=63, L=2869, F=13298)
WORKLOAD = "ba.py --bal_file data/fr1desk.txt (63 keyframes / 2869 landmarks / 13298 reprojection factors), defaults, 200 synchronous iterations"
OBS_PER_LMK = 10


def bench_config(world, synth_cams, synth_lmks):
    """The `config` object of the line: identical for both arms (ours / --impl reference) at the same N."""
    if world == 1:
        return {"workload": WORKLOAD, "step": "one 200-iteration solve from the initial state",
                "msgs_per_step": N_ITERS * 2 * FR1["F"], "parallelism": "1 GPU",
                "l2": "state reset + 512 MB L2 flush between timed steps; within a step the 10 MB state is L2-resident by nature"}
    F = synth_lmks * OBS_PER_LMK
    return {"workload": f"synthetic BAL {synth_cams} keyframes / {synth_lmks} landmarks / {F} reprojection factors (BASELINE configs[4]), "
                        f"landmark-partitioned over {world} GPUs, defaults, 200 synchronous iterations",
            "step": "one 200-iteration solve from the initial state", "msgs_per_step": N_ITERS * 2 * F,
            "parallelism": f"landmark partition over {world} GPUs, keyframes replicated, one exchange of keyframe partial sums per iteration",
            "l2": "5.8 GB of state streamed per iteration over all GPUs (inputs larger than L2; no flush needed)"}


_REAL_STDOUT = None


def protect_stdout():
    """The driver reads ONE JSON line from stdout, but libraries write there too (NCCL prints its version banner on
    stdout at NCCL_DEBUG=VERSION and WARN).  Keep a private handle on the real stdout and point fd 1 at stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_fr1desk():
    from gbp_b200 import balio
    G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
    return prob, G


def ncu_traffic(key):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]["bytes"]
    except Exception:
        return None


def b_alg(F, L, C, msg_cam_width=27):
    """Algorithmic bytes of one synchronous iteration and of the sweep kernel alone.  With the factored
    keyframe-message layout (18 instead of 27 doubles per message, read + written once) an edge moves 144 B less."""
    per_edge = 696 - (27 - msg_cam_width) * 8 * 2
    total = per_edge * F + 264 * L + 744 * C
    sweep = per_edge * F + 96 * L + 264 * C
    return total, sweep


def roofline_entry(workload, ms_per_launch, survey_bytes, layout_bytes, hbm_peak, traffic, msg_cam_width):
    """The `roofline` object of the line for the sweep kernel.  `achieved` / `frac` count the bytes this engine's layout
    has to move per launch (552 B per factor with factored keyframe messages) -- what the HBM really
    has to deliver, and what `traffic` (ncu dram bytes of the same launch) measures.  SURVEY 8(d)'s layout-independent
    figure (696 B per factor, full message rows) is kept as `achieved_survey_bytes` / `frac_survey_bytes`; it can exceed
    the peak because the kernel does not move those bytes."""
    sec = ms_per_launch * 1e-3
    moved = layout_bytes / sec / 1e9
    survey = survey_bytes / sec / 1e9
    return {"bound": "hbm", "kernel": "sweep_kernel", "workload": workload, "achieved": moved, "peak": hbm_peak, "unit": "GB/s",
            "frac": moved / hbm_peak, "algorithmic_bytes_per_launch": layout_bytes,
            "achieved_survey_bytes": survey, "frac_survey_bytes": survey / hbm_peak, "survey_bytes_per_launch": survey_bytes,
            "bytes_note": ("achieved / frac: bytes of the layout in use (" + ("552" if msg_cam_width != 27 else "696") + " B per factor + 96 B per landmark + 264 B per "
                           "keyframe, DESIGN.md section 3); *_survey_bytes: SURVEY 8(d)'s 696 B per factor (full 27-double keyframe message rows)"),
            "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read + dram__bytes_write per launch)",
            "ms_per_launch": ms_per_launch,
            "how": "CUDA events around every sweep_kernel launch on the engine's stream (gbp_ba_time_iterations, per_kernel=1)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = os.path.join(tempfile.gettempdir(), f"gbp_clocks_{os.getpid()}.csv")
        self.device = device
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.device), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


class Ctx:
    """torch / torch.distributed plumbing shared by the two workloads."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ.pop("NCCL_DEBUG")           # only a banner; anything NCCL still prints goes to stderr (protect_stdout)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        # a dedicated non-default stream shared by torch (events, NCCL ordering, L2 flush) and the engine:
        # the default stream's handle is 0, which the C ABI reads as "create your own stream"
        self.work_stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.work_stream)
        self.stream = self.work_stream.cuda_stream
        assert self.stream != 0
        self.hbm_peak, self.peak_src = peaks()
        self.flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def finish(self):
        if self.dist is not None:
            # never let a teardown problem hold the box: the numbers are out, leave within 30 s
            import threading
            threading.Timer(30.0, lambda: os._exit(0)).start()
            self.torch.cuda.synchronize()
            try:
                from gbp_b200.dist import shutdown
                shutdown()                  # the library's own NCCL communicator
                self.dist.destroy_process_group()
            except Exception as e:          # noqa: BLE001 -- the line is printed; a teardown problem must not turn the run into a failure
                sys.stderr.write(f"[bench] teardown: {type(e).__name__}: {e}\n")
            os._exit(0)


# ----------------------------------------------------------------------------------------------
def solve_200(graph):
    """The sweep schedule of ba.py:84-105 without the per-iteration client reads (device only)."""
    e = graph._eng
    e.iterate(3, robustify=True, local_relin=True)
    e.fill_iters(1)
    e.iterate(5, robustify=True, local_relin=True)
    e.fill_iters(1)
    e.iterate(N_ITERS - 8, robustify=True, local_relin=True)


def solve_api(graph):
    """The sweep schedule of ba.py:84-105 through the public graph API, no per-iteration host reads."""
    graph.iterate(3, robustify=True, local_relin=True)
    graph.reset_iters_since_relin(1)               # ba.py:91-93 at i = 3
    graph.iterate(5, robustify=True, local_relin=True)
    graph.reset_iters_since_relin(1)               # ... and at i = 8
    graph.iterate(N_ITERS - 8, robustify=True, local_relin=True)
    return graph.get_means()


def client_loop(graph):
    """The loop body of ba.py:84-105 through the public API (metrics + viewer reads every iteration)."""
    for i in range(N_ITERS):
        if i == 3 or i == 8:
            graph.reset_iters_since_relin(1)
        graph.metrics()                    # are(), energy(), relinearisation count  (ba.py:95-100)
        graph.cam_nodes[0].mu; graph.lmk_nodes[0].mu     # viewer.update reads the means (ba.py:103)
        graph.synchronous_iteration(robustify=True, local_relin=True)
    return graph.get_means()


def time_fr1desk(ctx, args, steps, warmup, with_e2e=True):
    """Device-timed and end-to-end 200-iteration fr1desk solves on this rank's GPU.  Returns a dict of raw results."""
    torch = ctx.torch
    from gbp_b200 import balio
    from gbp_b200.ba import create_ba_graph
    prob, G = load_fr1desk()
    tmp = tempfile.mkdtemp(prefix="gbp_bench_")
    bal_path = os.path.join(tmp, "fr1desk.txt")
    balio.write_bal(bal_path, prob, ["fr1desk (regenerated from tests/golden/fr1desk.npz, round-trip exact)"])
    graph = create_ba_graph(bal_path, CFG, device=ctx.local, stream=ctx.stream)
    eng = graph._eng
    F, Lm, C = eng.F, eng.L, eng.C
    mu_ref = np.concatenate([G["s199_cam_mu"], G["s199_lmk_mu"]])

    def prepare():
        graph.reset()
        graph.generate_priors_var(weaker_factor=CFG["prior_std_weaker_factor"])
        graph.update_all_beliefs()
        ctx.flush_buf.zero_()                     # L2 flush between timed steps (untimed)
        torch.cuda.synchronize()

    for _ in range(warmup):
        prepare(); solve_200(graph); torch.cuda.synchronize()
    step_ms = []
    launches0 = eng.launch_count()
    ctx.barrier()
    for _ in range(steps):
        prepare()
        e0, e1 = ctx.events()
        e0.record()
        solve_200(graph)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    ctx.barrier()
    out = {"F": F, "L": Lm, "C": C, "step_ms": step_ms, "launches": eng.launch_count() - launches0,
           "tile_edges": eng.tile_edges, "n_tiles": eng.n_tiles}
    mu = graph.get_means()
    out["parity_mu"] = float(np.max(np.abs(mu - mu_ref)) / np.max(np.abs(mu_ref)))
    out["are_final"] = graph.are()

    # per-kernel timing of the dominant kernel on the headline graph (separate, untimed for `value`)
    prepare()
    eng.iterate(20, True, True)
    tot_ms, sweep_ms = eng.time_iterations(100, True, True, per_kernel=True)
    _, sweep_b = b_alg(F, Lm, C)
    roof = {"bound": "hbm", "kernel": "sweep_kernel", "achieved": sweep_b / (sweep_ms / 100 * 1e-3) / 1e9, "peak": ctx.hbm_peak,
            "unit": "GB/s", "traffic": ncu_traffic("sweep_kernel/fr1desk"), "us_per_launch": sweep_ms / 100 * 1e3,
            "us_per_iteration_eager_with_events": tot_ms / 100 * 1e3,
            "note": "10 MB working set is L2-resident: latency/launch-bound, HBM fraction is not meaningful here"}
    roof["frac"] = roof["achieved"] / ctx.hbm_peak
    out["roofline_fr1desk"] = roof
    if not with_e2e:
        graph.close()
        return out

    # ------------------------------------------------------------------ e2e through the public API
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy()
              for k, v in dict(cam_id=prob.cam_id, lmk_id=prob.lmk_id, z=prob.z, cam=prob.cam_means, lmk=prob.lmk_means).items()}
    snap_bytes = eng.snapshot_layout()[0]
    # bytes the engine really moves: the compiled graph (slot-ordered ids + measurements, tiles, both CSR tables) and
    # the initial means go up; every snapshot (metrics + compact means) comes down
    out["h2d"] = eng.n_slots * (4 + 16) + eng.n_tiles * (8 + 4) + F * (4 + 4) + (Lm + 1) * 4 + (C + 1) * 4 + (6 * C + 3 * Lm) * 8
    out["snap_bytes"] = snap_bytes

    def e2e_run(loop, tag):
        times, phases, means = [], [], None
        for it in range(warmup + steps):
            gc.collect()                      # graphs of earlier steps (proxy objects, pinned blocks) die here, untimed
            ctx.flush_buf.zero_(); ctx.barrier()
            gc.disable()                      # like timeit: no collector pause inside the timed region
            t0 = time.perf_counter()
            p2 = balio.BALProblem(pinned["cam_id"], pinned["lmk_id"], pinned["z"], pinned["cam"], pinned["lmk"], prob.K4)
            g2 = create_ba_graph(p2, CFG, device=ctx.local, stream=ctx.stream)
            t1 = time.perf_counter()
            g2.generate_priors_var(weaker_factor=CFG["prior_std_weaker_factor"])
            g2.update_all_beliefs()
            t2 = time.perf_counter()
            means = loop(g2)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            gc.enable()
            if os.environ.get("GBP_BENCH_DEBUG"):
                print(f"[{tag} {it}] create {1e3 * (t1 - t0):.2f} ms  priors {1e3 * (t2 - t1):.2f}  loop {1e3 * (t3 - t2):.2f}  total {1e3 * (t3 - t0):.2f}", file=sys.stderr)
            g2.close()
            if it >= warmup:
                times.append(t3 - t0)
                phases.append((t1 - t0, t2 - t1, t3 - t2))
        ph = 1e3 * np.array(phases)
        return {"t": ctx.max_over_ranks(float(np.sum(times))), "parity": float(np.max(np.abs(means - mu_ref)) / np.max(np.abs(mu_ref))),
                "phases_ms": {"create_ba_graph": {"p50": float(np.median(ph[:, 0])), "max": float(ph[:, 0].max())},
                              "priors_and_first_beliefs": {"p50": float(np.median(ph[:, 1])), "max": float(ph[:, 1].max())},
                              "iterations_and_readback": {"p50": float(np.median(ph[:, 2])), "max": float(ph[:, 2].max())},
                              "step": {"p50": float(1e3 * np.median(times)), "max": float(1e3 * np.max(times)), "min": float(1e3 * np.min(times))}}}

    out["e2e"] = e2e_run(solve_api, "e2e")
    out["e2e_loop"] = e2e_run(client_loop, "e2e_client_loop")
    graph.close()
    return out


def run_unmodified_ba_py():
    """Wall time of the reference's OWN ba.py (staged copy, baseline/_ref) against this engine, 200 iterations on fr1desk."""
    script = os.path.join(REF_COPY, "ba.py")
    if not os.path.exists(script):
        return {"unavailable": "baseline/_ref not staged (__graft_entry__.build() does it where /root/reference exists)"}
    t0 = time.perf_counter()
    try:
        res = subprocess.run([sys.executable, "-m", "gbp_b200.run", script, "--bal_file", "data/fr1desk.txt"], cwd=REF_COPY,
                             env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=300)
    except Exception as e:       # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    wall = time.perf_counter() - t0
    last = [l for l in res.stdout.splitlines() if l.startswith("Iteration")]
    _, G = load_fr1desk()
    ref_line = f"Iteration 199 // ARE {G['are'][199]:.4f} // Energy {G['energy'][199]:.4f} // Num factors relinearising {int(G['n_relin'][199])}"
    return {"rc": res.returncode, "wall_s": wall, "iterations_printed": len(last), "last_line": last[-1] if last else None,
            "reference_last_line_from_fixture": ref_line,
            "what": "python -m gbp_b200.run baseline/_ref/ba.py --bal_file data/fr1desk.txt: interpreter start, BAL parse, graph build, 200 x (are, energy, Python loop over 13298 factor proxies, viewer read, sweep)"}


def bench_ours(args):
    protect_stdout()
    ctx = Ctx()
    if ctx.world > 1:
        return bench_partitioned(ctx, args)
    torch = ctx.torch
    sampler = ClockSampler(ctx.local)
    sampler.start()
    r = time_fr1desk(ctx, args, args.steps, args.warmup)
    F = r["F"]
    msgs_per_step = N_ITERS * 2 * F
    t_step = float(np.sum(r["step_ms"])) / 1e3
    value = args.steps * msgs_per_step / t_step
    ms_per_step = 1e3 * t_step / args.steps
    e2e, loop = r["e2e"], r["e2e_loop"]

    # ------------------------------------------------------------------ synthetic 10M-factor graph
    synth, roofline = None, r["roofline_fr1desk"]
    if not args.no_synthetic:
        synth, roofline = bench_synthetic_1gpu(ctx, args)
    clocks = sampler.stop()
    ba_py = None if args.no_ba_py else run_unmodified_ba_py()
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample(args, ctx)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "fr1desk measurements from the committed fixture (TUM RGB-D / ORB-SLAM keyframes); synthetic BAL graph for the roofline",
        "config": bench_config(1, args.synth_cams, args.synth_lmks),
        "engine": {"tile_edges": r["tile_edges"], "n_tiles": r["n_tiles"]},
        "us_per_iteration": 1e3 * ms_per_step / N_ITERS,
        "clocks": clocks,
        "e2e": {"value": args.steps * msgs_per_step / e2e["t"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 3 * r["snap_bytes"],
                "ms_per_step": 1e3 * e2e["t"] / args.steps, "phases_ms": e2e["phases_ms"],
                "what": "create_ba_graph from pinned host arrays (graph compile + upload) + priors + 200 synchronous iterations with the resets at 3 and 8 + final means on the host; wall clock"},
        "e2e_client_loop": {"value": args.steps * msgs_per_step / loop["t"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                            "d2h_bytes_per_step": (N_ITERS + 1) * r["snap_bytes"], "ms_per_step": 1e3 * loop["t"] / args.steps,
                            "max_rel_err": loop["parity"], "phases_ms": loop["phases_ms"],
                            "what": "the same solve driven like ba.py's loop body: 200 x (are, energy, relinearisation count and the viewer's means read back to the host, synchronous_iteration); a serial CPU<->GPU ping-pong"},
        "unmodified_ba_py": ba_py,
        "gpu_launches": r["launches"],
        "parity": {"max_rel_err_means_vs_reference_fixture": r["parity_mu"], "e2e_max_rel_err": e2e["parity"], "tol": 1e-4,
                   "ok": bool(r["parity_mu"] < 1e-4 and e2e["parity"] < 1e-4 and loop["parity"] < 1e-4), "final_are_px": r["are_final"]},
        "roofline": roofline, "roofline_fr1desk": r["roofline_fr1desk"], "peak_source": ctx.peak_src,
        "cpu_baseline": cpu, "synthetic": synth,
    }
    emit(json.dumps(line))
    ctx.finish()


def synth_info(eng, F, Lm, C):
    total_b, _ = b_alg(F, Lm, C)
    total_b_layout, _ = b_alg(F, Lm, C, eng.msg_cam_width)
    return total_b, total_b_layout, {
        "msg_cam_doubles": eng.msg_cam_width, "sweep_kernel_build": eng.sweep_variant, "l2_prefetch_tiles": eng.prefetch_tiles,
        "tile_edges": eng.tile_edges, "n_tiles_local": eng.n_tiles,
        "algorithmic_bytes_per_iteration_this_layout": total_b_layout,
        "note": "factor->keyframe messages are stored with their rank-2 precision factored (18 doubles instead of 27): 144 B per factor less than SURVEY 8(d)'s 696 B"}


def bench_synthetic_1gpu(ctx, args):
    """configs[3]: 1k keyframes / 1M landmarks / 10M factors on one GPU; roofline of the sweep kernel."""
    torch = ctx.torch
    from gbp_b200.synthetic import make_synthetic
    from gbp_b200.dist import PartitionedBAGraph
    t0 = time.perf_counter()
    prob = make_synthetic(args.synth_cams, args.synth_lmks, OBS_PER_LMK, seed=0)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    pg = PartitionedBAGraph(prob, CFG, device=ctx.local, stream=ctx.stream)
    build_s = time.perf_counter() - t0
    pg.generate_priors_var(CFG["prior_std_weaker_factor"])
    pg.update_all_beliefs()
    F, Lm, C = prob.n_edges, prob.n_points, prob.n_keyframes
    k = args.synth_iters
    for _ in range(max(args.warmup, 3)):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    torch.cuda.synchronize()
    l0 = pg.engine.launch_count()
    e0, e1 = ctx.events()
    e0.record()
    for _ in range(k):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    e1.record()
    torch.cuda.synchronize()
    launches = pg.engine.launch_count() - l0
    t = e0.elapsed_time(e1) / 1e3
    eng = pg.engine
    total_b, total_b_layout, layout = synth_info(eng, F, Lm, C)
    _, sweep_b_survey = b_alg(F, Lm, C)
    _, sweep_b_local = b_alg(F, Lm, C, eng.msg_cam_width)
    tot_ms, sweep_ms = eng.time_iterations(k, True, True, per_kernel=True)
    # sustained: 200 back-to-back iterations (~0.3 s of continuous fp64 + HBM load; the burst above is ~30 ms)
    ks = args.synth_sustained
    s0, s1 = ctx.events()
    s0.record()
    for _ in range(ks):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    s1.record()
    torch.cuda.synchronize()
    t_sus = s0.elapsed_time(s1) / 1e3
    are, energy, _ = pg.metrics()
    synth = {"workload": f"synthetic BAL {C} keyframes / {Lm} landmarks / {F} factors, 1 GPU",
             "value": k * 2 * F / t, "unit": UNIT, "ms_per_iteration": 1e3 * t / k, "iterations_timed": k,
             "algorithmic_bytes_per_iteration": total_b_layout, "survey_bytes_per_iteration": total_b,
             "achieved_gbs_whole_iteration": total_b_layout / (t / k) / 1e9,
             "frac_of_hbm_peak_whole_iteration": total_b_layout / (t / k) / 1e9 / ctx.hbm_peak,
             "frac_of_hbm_peak_whole_iteration_survey_bytes": total_b / (t / k) / 1e9 / ctx.hbm_peak,
             "sustained": {"iterations": ks, "ms_per_iteration": 1e3 * t_sus / ks, "value": ks * 2 * F / t_sus, "unit": UNIT,
                           "frac_of_hbm_peak_whole_iteration": total_b_layout / (t_sus / ks) / 1e9 / ctx.hbm_peak,
                           "note": "back-to-back iterations for ~0.3 s; the burst figure above times %d iterations" % k},
             "l2": "%.1f GB streamed per iteration (inputs larger than L2, no flush needed)" % (total_b_layout / 1e9),
             "gpu_launches": launches, "n_iterations_applied": pg.n_iterations, "are_px_after": are, "energy_after": energy,
             "generate_s": gen_s, "graph_build_s": build_s, "ms_per_iteration_eager_with_events": tot_ms / k, "layout": layout}
    roof = roofline_entry(synth["workload"], sweep_ms / k, sweep_b_survey, sweep_b_local, ctx.hbm_peak,
                          ncu_traffic(f"sweep_kernel/synthetic_{C}_{Lm}_{F}"), eng.msg_cam_width)
    pg.close()
    return synth, roof


def L_nccl_version():
    from gbp_b200 import _lib as L
    return L.comm_version()


def phase_breakdown(ctx, pg, n=30):
    """Where an iteration of the partitioned graph spends its time on this rank: the three phases issued one after the other on
    the engine's stream with CUDA events between them (medians over n iterations, max over ranks).  In the real iteration (one
    CUDA-graph replay) the exchange runs on a high-priority side stream WHILE the landmark beliefs are updated, so the captured
    iteration is shorter than the sum of the phases."""
    from gbp_b200 import _lib as L
    torch = ctx.torch
    eng = pg.engine
    st = L.ST_MESSAGES | L.ST_BELIEFS | L.ST_DEFER_LANDMARKS | L.ST_ROBUSTIFY | L.ST_RELIN | L.ST_LOCAL_DAMPING
    rows = []
    for it in range(n + 5):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        eng.sweep_local(st)                         # sweep_kernel + keyframe chunk sums (belief_kernel, keyframe CTAs)
        ev[1].record()
        eng.exchange()                              # ncclAllGather of the chunk sums + cam_update_kernel
        ev[2].record()
        eng.landmark_update()                       # belief_kernel, landmark CTAs
        ev[3].record()
        torch.cuda.synchronize()
        pg.n_iterations += 1
        if it >= 5:
            rows.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)] + [ev[0].elapsed_time(ev[3]) * 1e3])
    med = np.median(np.array(rows), axis=0)
    med = [ctx.max_over_ranks(float(x)) for x in med]
    names = ["local_sweep_and_keyframe_chunk_sums", "exchange_and_keyframe_update", "landmark_belief_update", "total_serialised"]
    out = dict(zip(names, med))
    out["note"] = "phases serialised on one stream with CUDA events between them, median of %d iterations, max over ranks; the timed solve replays one CUDA graph per iteration in which the exchange (side stream) overlaps the landmark update" % n
    return out


def bench_partitioned(ctx, args):
    """N > 1: configs[4], the synthetic graph landmark-partitioned over the ranks; headline of the line (strong scaling)."""
    torch, dist, rank, world = ctx.torch, ctx.dist, ctx.rank, ctx.world
    from gbp_b200.synthetic import make_synthetic
    from gbp_b200.dist import PartitionedBAGraph
    from gbp_b200.balio import BALProblem
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    prob = make_synthetic(args.synth_cams, args.synth_lmks, OBS_PER_LMK, seed=0)
    gen_s = time.perf_counter() - t0
    F, Lm, C = prob.n_edges, prob.n_points, prob.n_keyframes
    S = N_ITERS

    def build():
        g = PartitionedBAGraph(prob, CFG, rank=rank, world=world, device=ctx.local, dist=dist, torch_stream=ctx.work_stream)
        return g

    def prepare(g):
        g.generate_priors_var(CFG["prior_std_weaker_factor"])
        g.update_all_beliefs()

    def solve(g):
        g.iterate(S, robustify=True, local_relin=True)

    t0 = time.perf_counter()
    pg = build()
    build_s = time.perf_counter() - t0
    prepare(pg)
    # launches of one (eager) iteration, counted by the engine; the captured replays launch the same kernels
    l0 = pg.engine.launch_count()
    pg.synchronous_iteration(robustify=True, local_relin=True)
    per_iter_launches = pg.engine.launch_count() - l0
    captured = pg.capture(local_relin=True, robustify=True) if not args.no_capture else False
    for _ in range(args.warmup):
        pg.reset(); prepare(pg); solve(pg)
    step_ms = []
    ctx.barrier()
    for _ in range(args.steps):
        pg.reset(); prepare(pg)
        ctx.barrier()
        e0, e1 = ctx.events()
        e0.record()
        solve(pg)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    t_dev = ctx.max_over_ranks(float(np.sum(step_ms)) / 1e3)
    ctx.barrier()
    n_applied = pg.n_iterations
    are, energy, nrel = pg.metrics()           # the state after exactly S iterations (the breakdown below sweeps on)
    means = pg.get_means()                     # collective: every rank
    breakdown = phase_breakdown(ctx, pg)
    eng = pg.engine
    total_b, total_b_layout, layout = synth_info(eng, F, Lm, C)
    snap_local = (6 * C + 3 * eng.L) * 8
    h2d_local = eng.n_slots * (4 + 16) + eng.n_tiles * (8 + 4) + eng.F * (4 + 4) + (eng.L + 1) * 4 + (C + 1) * 4 + (6 * C + 3 * eng.L) * 8
    pg.close()

    # ---- end to end: from the host arrays of the whole problem to this rank's means on the host, every step
    e2e_times = []
    for it in range(args.warmup + args.steps):
        gc.collect()
        ctx.barrier()
        t0 = time.perf_counter()
        g2 = build()
        prepare(g2)
        if not args.no_capture:
            g2.capture(local_relin=True, robustify=True)
        solve(g2)
        g2.adapter.cam_means(); g2.adapter.lmk_means()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        g2.close()
        if it >= args.warmup:
            e2e_times.append(dt)
    t_e2e = ctx.max_over_ranks(float(np.sum(e2e_times)))
    h2d = ctx.max_over_ranks(float(h2d_local)) * world
    ctx.barrier()

    # ---- rank 0: the same solve on ONE GPU in the same process (parity of the partitioned path + 1-GPU time on this box)
    par = None
    if rank == 0 and not args.no_parity_1gpu:
        from gbp_b200.engine import BAEngine
        from gbp_b200 import _lib as L
        e1g = BAEngine(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG, device=ctx.local, stream=ctx.stream)
        e1g.generate_priors(CFG["prior_std_weaker_factor"], e1g.prior_scan())
        e1g.update_beliefs()
        a0, a1 = ctx.events()
        a0.record()
        e1g.iterate(S, robustify=True, local_relin=True)
        a1.record()
        torch.cuda.synchronize()
        ms_1gpu = a0.elapsed_time(a1)
        m1 = e1g.metrics()
        mu1 = np.concatenate([e1g.read(L.F_CAM_MU).ravel(), e1g.read(L.F_LMK_MU).ravel()])
        e1g.close()
        par = {"n_iters": S, "n_iterations_applied_ngpu": n_applied, "are_1gpu": float(m1[0]) / F, "are_ngpu": are,
               "energy_1gpu": float(m1[1]), "energy_ngpu": energy,
               "max_rel_err_means": float(np.max(np.abs(means - mu1)) / np.max(np.abs(mu1))),
               "max_abs_err_means": float(np.max(np.abs(means - mu1))),
               "single_gpu_ms_per_iteration_same_box": ms_1gpu / S,
               "speedup_vs_1gpu_same_box": (ms_1gpu / S) / (1e3 * t_dev / (args.steps * S)),
               "what": "rank 0 runs the single-GPU engine on the WHOLE graph for the same 200 iterations from the same initial state, in this process"}
    ctx.barrier()

    # ---- the replicated fr1desk solve (does not shard: every rank solves the same 10 MB graph)
    fr1 = None
    if not args.no_fr1desk_replicas:
        r = time_fr1desk(ctx, args, 3, 3, with_e2e=False)
        t = ctx.max_over_ranks(float(np.sum(r["step_ms"])) / 1e3)
        fr1 = {"value_per_replica": 3 * N_ITERS * 2 * r["F"] / t, "unit": UNIT, "replicas": world, "us_per_iteration": 1e6 * t / (3 * N_ITERS),
               "max_rel_err_means_vs_reference_fixture": r["parity_mu"],
               "note": "N independent replicas of the 200-iteration fr1desk solve (a 10 MB graph does not shard); per-replica throughput, NOT multiplied by N"}
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        msgs_per_step = S * 2 * F
        line = {
            "metric": METRIC, "value": args.steps * msgs_per_step / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic BAL graph (gbp_b200/synthetic.py, seed 0): keyframes on a circle, landmarks in a cube, 10 observations per landmark, 2 px noise",
            "config": bench_config(world, args.synth_cams, args.synth_lmks),
            "engine": {"exchange": "ncclAllGather of the keyframe chunk sums, called by libgbp_b200 (gbp_ba_attach_comm) on a high-priority side stream; overlaps the landmark belief update", "nccl_version": L_nccl_version(), "iteration_captured_in_cuda_graph": bool(captured), "layout": layout},
            "ms_per_iteration": 1e3 * t_dev / (args.steps * S), "clocks": clocks,
            "algorithmic_bytes_per_iteration": total_b_layout,
            "frac_of_hbm_peak_whole_iteration_per_gpu": total_b_layout / world / (t_dev / (args.steps * S)) / 1e9 / ctx.hbm_peak,
            "e2e": {"value": args.steps * msgs_per_step / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (6 * C * world + 3 * Lm) * 8,
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "what": "per rank, from the host arrays of the whole problem: cut out the local landmark block, compile + upload the local graph, priors (cross-rank max), capture, 200 synchronous iterations, local means back on the host; wall clock, max over ranks"},
            "gpu_launches": per_iter_launches * S * args.steps,
            "gpu_launches_note": f"{per_iter_launches} kernels of this library per iteration and rank (sweep, keyframe partial sums, landmark beliefs, keyframe update) + the exchange; counted on rank 0",
            "iteration_phases_us": breakdown,
            "parity_vs_1gpu": par, "are_px_after": are, "energy_after": energy,
            "generate_s": gen_s, "graph_build_s": build_s, "fr1desk_replicas": fr1, "peak_source": ctx.peak_src,
            "roofline": None, "cpu_baseline": None,
        }
        emit(json.dumps(line))
    ctx.finish()


# ----------------------------------------------------------------------------------------------
def _c_oracle_solves(n_solves, threads):
    """n_solves complete 200-iteration fr1desk solves by the plain-C OpenMP port of the reference algorithm."""
    from oracle import c_oracle
    c_oracle.set_threads(threads)
    prob, G = load_fr1desk()
    times, o = [], None
    for _ in range(n_solves):
        o = c_oracle.COracle(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG)
        o.generate_priors_var(CFG["prior_std_weaker_factor"])
        o.update_all_beliefs()
        t0 = time.perf_counter()
        for i in range(N_ITERS):
            if i == 3 or i == 8:
                o.fill_iters(1)
            o.synchronous_iteration(robustify=True, local_relin=True)
        times.append(time.perf_counter() - t0)
    mu_ref = np.concatenate([G["s199_cam_mu"], G["s199_lmk_mu"]])
    mu = np.concatenate([o.cam_mu.ravel(), o.lmk_mu.ravel()])
    return times, o, float(np.max(np.abs(mu - mu_ref)) / np.max(np.abs(mu_ref)))


def time_reference_numpy(n_sweeps):
    """The UNMODIFIED reference (staged copy under baseline/_ref) timed on this host: subprocess with cwd = baseline/_ref
    so that its own packages (gbp, utils) are the ones imported; one thread (CPU-baseline plan)."""
    if not os.path.exists(os.path.join(REF_COPY, "gbp", "gbp_ba.py")):
        return {"unavailable": "baseline/_ref not staged (__graft_entry__.build() does it where /root/reference exists)"}
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    env.pop("PYTHONPATH", None)
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "time_reference.py"), "data/fr1desk.txt", str(n_sweeps)],
                             cwd=REF_COPY, env=env, capture_output=True, text=True, timeout=600)
        d = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    except Exception as e:       # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    return {"value": d["msgs_per_s"], "unit": UNIT, "cores": 1, "kind": "reference", "host_cpus": os.cpu_count(),
            "median_iteration_s": d["median_iteration_s"], "iteration_s": d["iteration_s"], "create_ba_graph_s": d["create_s"],
            "are_after": d["are_after"], "n_sweeps": n_sweeps,
            "sample": f"unmodified reference (baseline/_ref, imported from {d['module']}): create_ba_graph + priors + {n_sweeps} x "
                      f"synchronous_iteration(robustify=True, local_relin=True) on data/fr1desk.txt, median per iteration; NumPy {d['numpy']}, 1 thread"}


def cpu_baseline_sample(args, ctx=None):
    """CPU baseline on the GPU box's host cores: the plain-C OpenMP port (oracle/gbp_oracle.c) on all threads = the stiffer
    arm (`kind: port`), and the UNMODIFIED NumPy reference timed in the same run (`reference_numpy`, `kind: reference`)."""
    threads = os.cpu_count() or 1
    times, o, err = _c_oracle_solves(12, threads)
    times = times[2:]
    c_val = len(times) * N_ITERS * 2 * o.F / float(np.sum(times))
    ref = time_reference_numpy(args.ref_sweeps)
    if ctx is not None and "are_after" in ref:
        # live parity against the reference run of this very bench: same file, same number of sweeps, no client resets
        from gbp_b200.ba import create_ba_graph
        g = create_ba_graph(os.path.join(REF_COPY, "data", "fr1desk.txt"), CFG, device=ctx.local, stream=ctx.stream)
        g.generate_priors_var(weaker_factor=CFG["prior_std_weaker_factor"])
        g.update_all_beliefs()
        g.iterate(args.ref_sweeps, robustify=True, local_relin=True)
        ours = g.are()
        g.close()
        ref["are_after_ours_same_sweeps"] = ours
        ref["are_rel_diff"] = abs(ours - ref["are_after"]) / abs(ref["are_after"])
    return {"value": c_val, "unit": UNIT, "cores": o.threads, "kind": "port",
            "sample": f"{len(times)} complete 200-iteration fr1desk solves by oracle/gbp_oracle.c (plain C + OpenMP, reference arithmetic form), {float(np.sum(times)):.1f} s; means {err:.1e} from the reference fixture",
            "host_cpus": os.cpu_count(), "reference_numpy": ref}


def bench_reference(args):
    """Reference arm: the reference's algorithm on the host cores -- the plain-C OpenMP port of it (oracle/gbp_oracle.c,
    pinned against the reference's fixtures) with every host thread: the STIFFER CPU arm (the unmodified NumPy reference is
    ~2000x slower and is timed under cpu_baseline.reference_numpy of the N = 1 line).  N = 1: one complete 200-iteration
    fr1desk solve per step.  N > 1 (rank 0 only): the workload of our N-GPU arm is the 10 M-factor synthetic graph; a step
    is a bounded sample of it -- synchronous iterations of a 1/10-scale instance of the same generator."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    threads = os.cpu_count() or 1
    cfg = bench_config(world, args.synth_cams, args.synth_lmks)
    if world == 1:
        times, o, err = _c_oracle_solves(args.warmup + args.steps, threads)
        times = times[args.warmup:]
        msgs = N_ITERS * 2 * o.F
        value = args.steps * msgs / float(np.sum(times))
        sample = f"one complete 200-iteration fr1desk solve per step by oracle/gbp_oracle.c (plain C + OpenMP port of the reference algorithm, {o.threads} threads)"
        data = "fr1desk measurements from the committed fixture"
        nthreads = o.threads
    else:
        from oracle import c_oracle
        from gbp_b200.synthetic import make_synthetic
        c_oracle.set_threads(threads)
        prob = make_synthetic(max(args.synth_cams // 10, 10), max(args.synth_lmks // 10, 1000), OBS_PER_LMK, seed=0)
        o = c_oracle.COracle(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG)
        o.generate_priors_var(CFG["prior_std_weaker_factor"])
        o.update_all_beliefs()
        per_step = 5
        times = []
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for _ in range(per_step):
                o.synchronous_iteration(robustify=True, local_relin=True)
            times.append(time.perf_counter() - t0)
        times = times[args.warmup:]
        value = args.steps * per_step * 2 * o.F / float(np.sum(times))
        err = None
        sample = (f"{per_step} synchronous iterations per step of a 1/10-scale instance of the same generator ({prob.n_keyframes} keyframes / {prob.n_points} landmarks / "
                  f"{o.F} factors) by oracle/gbp_oracle.c ({o.threads} threads); msgs/s is size-independent to first order")
        data = "synthetic BAL graph (gbp_b200/synthetic.py, seed 0), 1/10 scale"
        nthreads = o.threads
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": data, "config": cfg,
            "parity": {"max_rel_err_means_vs_reference_fixture": err},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample, "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-synthetic", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba-py", action="store_true", help="skip the subprocess run of the unmodified ba.py")
    ap.add_argument("--no-capture", action="store_true", help="multi-GPU: do not capture the iteration in a CUDA graph")
    ap.add_argument("--no-parity-1gpu", action="store_true", help="multi-GPU: skip the single-GPU solve of the whole graph on rank 0")
    ap.add_argument("--no-fr1desk-replicas", action="store_true")
    ap.add_argument("--synth-cams", type=int, default=1000)
    ap.add_argument("--synth-lmks", type=int, default=1_000_000)
    ap.add_argument("--synth-iters", type=int, default=20)
    ap.add_argument("--synth-sustained", type=int, default=200)
    ap.add_argument("--ref-sweeps", type=int, default=5, help="iterations of the unmodified NumPy reference timed for cpu_baseline.reference_numpy")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        bench_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3          # timing rule: at least 3 warm-up steps
        bench_ours(args)


if __name__ == "__main__":
    main()
