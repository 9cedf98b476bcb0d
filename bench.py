#!/usr/bin/env python
"""Benchmark of the GBP bundle-adjustment sweep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE configs[2]): ``ba.py --bal_file data/fr1desk.txt``, defaults, 200
synchronous iterations.  One STEP = one complete 200-iteration solve from the initial state
(including the client's `iters_since_relin = 1` resets at iterations 3 and 8, ba.py:91-93);
the graph state is reset and L2 is flushed between steps (both untimed), inside a step the
10 MB working set is legitimately L2-resident, as in a real run.
  value  = GBP messages/s = 200 * 2F / device time of the step (CUDA events on the engine's stream)
  e2e    = the same step through the public Python API starting from pinned HOST arrays: create_ba_graph
           (graph compile + upload), priors, the 200 synchronous iterations with the two resets, final
           means read back to the host; wall clock with synchronisation.  `e2e_client_loop` is the same
           solve driven exactly like ba.py's loop body (are / energy / relinearisation count and the
           viewer's means read back to the host between every two sweeps).
The same run also measures the synthetic 1k-camera / 1M-landmark / 10M-factor graph (configs[3];
7.2 GB of state streamed per iteration, far larger than L2), which is where the HBM roofline is
meaningful: `roofline` refers to the sweep kernel on that graph, `roofline_fr1desk` to the
(latency-bound, L2-resident) headline graph.

``--impl reference`` times the CPU restatement of the reference algorithm (oracle/gbp_oracle.c: plain C +
OpenMP on every host thread; the Python reference itself cannot travel to the GPU box) on the same workload.
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

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8,
           eta_damping=0.4, prior_std_weaker_factor=50.0)
N_ITERS = 200
METRIC = "gbp_messages_per_sec"
UNIT = "msgs/s"
WORKLOAD = "ba.py --bal_file data/fr1desk.txt (63 keyframes / 2869 landmarks / 13298 reprojection factors), defaults, 200 synchronous iterations"


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
    """The `roofline` object of the bench line for the sweep kernel.  `achieved` follows the contract: SURVEY 8(d)'s
    algorithmic bytes (696 B per factor) / mean launch duration.  With factored keyframe messages the kernel has to move
    only 552 B per factor, so that figure can exceed the peak; `achieved_moved` / `frac_moved` count the bytes of the
    layout in use (what the HBM really has to deliver; `traffic` is the ncu measurement of the same)."""
    sec = ms_per_launch * 1e-3
    ach = survey_bytes / sec / 1e9
    moved = layout_bytes / sec / 1e9
    return {"bound": "hbm", "kernel": "sweep_kernel", "workload": workload, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach / hbm_peak, "algorithmic_bytes_per_launch": survey_bytes,
            "achieved_moved": moved, "frac_moved": moved / hbm_peak, "moved_bytes_per_launch": layout_bytes,
            "bytes_note": ("achieved / frac use SURVEY 8(d)'s algorithmic bytes (696 B per factor).  "
                           + ("This engine stores factor->keyframe messages factored (18 instead of 27 doubles) and has to move only "
                              "552 B per factor, so achieved may exceed the HBM peak; achieved_moved / frac_moved count those bytes "
                              "and are the distance to the HBM roofline, traffic is their ncu measurement."
                              if msg_cam_width != 27 else "Full message rows: the layout moves exactly those bytes.")),
            "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full, per launch)",
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


def bench_ours(args):
    protect_stdout()
    import torch
    from gbp_b200 import balio
    from gbp_b200.ba import create_ba_graph

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")           # only a banner; anything NCCL still prints goes to stderr (protect_stdout)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated non-default stream shared by torch (events, NCCL ordering, L2 flush) and the engine:
    # the default stream's handle is 0, which the C ABI reads as "create your own stream"
    work_stream = torch.cuda.Stream()
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0
    hbm_peak, peak_src = peaks()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ headline: fr1desk
    prob, G = load_fr1desk()
    tmp = tempfile.mkdtemp(prefix="gbp_bench_")
    bal_path = os.path.join(tmp, "fr1desk.txt")
    balio.write_bal(bal_path, prob, ["fr1desk (regenerated from tests/golden/fr1desk.npz, round-trip exact)"])
    graph = create_ba_graph(bal_path, CFG, device=local, stream=stream)
    eng = graph._eng
    F, Lm, C = eng.F, eng.L, eng.C
    msgs_per_step = N_ITERS * 2 * F

    def prepare():
        graph.reset()
        graph.generate_priors_var(weaker_factor=CFG["prior_std_weaker_factor"])
        graph.update_all_beliefs()
        flush_buf.zero_()                     # L2 flush between timed steps (untimed)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        prepare(); solve_200(graph); torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms = []
    launches0 = eng.launch_count()
    barrier()
    for _ in range(args.steps):
        prepare()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solve_200(graph)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = eng.launch_count() - launches0
    # parity of the last timed solve against the fixture generated from the reference
    mu = graph.get_means()
    mu_ref = np.concatenate([G["s199_cam_mu"], G["s199_lmk_mu"]])
    parity_mu = float(np.max(np.abs(mu - mu_ref)) / np.max(np.abs(mu_ref)))
    are_final = graph.are()
    t_step = max_over_ranks(float(np.sum(step_ms)) / 1e3)          # seconds for K steps, max over ranks
    value = world * args.steps * msgs_per_step / t_step
    ms_per_step = 1e3 * t_step / args.steps

    # per-kernel timing of the dominant kernel on the headline graph (separate, untimed for `value`)
    prepare()
    eng.iterate(20, True, True)
    tot_ms, sweep_ms = eng.time_iterations(100, True, True, per_kernel=True)
    total_b, sweep_b = b_alg(F, Lm, C)
    roof_fr1 = {"bound": "hbm", "kernel": "sweep_kernel", "achieved": sweep_b / (sweep_ms / 100 * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "traffic": ncu_traffic("sweep_kernel/fr1desk"), "us_per_launch": sweep_ms / 100 * 1e3, "us_per_iteration": tot_ms / 100 * 1e3,
                "note": "10 MB working set is L2-resident: latency/launch-bound, HBM fraction is not meaningful here"}
    roof_fr1["frac"] = roof_fr1["achieved"] / hbm_peak

    # ------------------------------------------------------------------ e2e through the public API
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy()
              for k, v in dict(cam_id=prob.cam_id, lmk_id=prob.lmk_id, z=prob.z, cam=prob.cam_means, lmk=prob.lmk_means).items()}
    snap_bytes = eng.snapshot_layout()[0]
    # bytes the engine really moves: the compiled graph (slot-ordered ids + measurements, tiles, both CSR tables) and
    # the initial means go up; every snapshot (metrics + compact means) comes down
    h2d = eng.n_slots * (4 + 16) + eng.n_tiles * (8 + 4) + F * (4 + 4) + (Lm + 1) * 4 + (C + 1) * 4 + (6 * C + 3 * Lm) * 8

    def e2e_run(loop, tag):
        times, means = [], None
        for it in range(args.warmup + args.steps):
            gc.collect()                      # graphs of earlier steps (proxy objects, pinned blocks) die here, untimed
            flush_buf.zero_(); barrier()
            gc.disable()                      # like timeit: no collector pause inside the timed region
            t0 = time.perf_counter()
            p2 = balio.BALProblem(pinned["cam_id"], pinned["lmk_id"], pinned["z"], pinned["cam"], pinned["lmk"], prob.K4)
            g2 = create_ba_graph(p2, CFG, device=local, stream=stream)
            t1 = time.perf_counter()
            g2.generate_priors_var(weaker_factor=CFG["prior_std_weaker_factor"])
            g2.update_all_beliefs()
            t2 = time.perf_counter()
            means = loop(g2)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            gc.enable()
            if os.environ.get("GBP_BENCH_DEBUG"):
                print(f"[{tag} {it}] create {1e3 * (t1 - t0):.2f} ms  priors {1e3 * (t2 - t1):.2f}  loop {1e3 * (t0 + dt - t2):.2f}  total {1e3 * dt:.2f}", file=sys.stderr)
            g2.close()
            if it >= args.warmup:
                times.append(dt)
        t = max_over_ranks(float(np.sum(times)))
        return t, float(np.max(np.abs(means - mu_ref)) / np.max(np.abs(mu_ref)))

    e2e_t, e2e_parity = e2e_run(solve_api, "e2e")
    e2e_val = world * args.steps * msgs_per_step / e2e_t
    d2h = 3 * snap_bytes                             # one snapshot behind each of the three iterate() calls
    loop_t, loop_parity = e2e_run(client_loop, "e2e_client_loop")
    loop_val = world * args.steps * msgs_per_step / loop_t
    d2h_loop = (N_ITERS + 1) * snap_bytes
    graph.close()

    # ------------------------------------------------------------------ synthetic 10M-factor graph
    synth = None
    roofline = roof_fr1
    if not args.no_synthetic:
        synth, roofline = bench_synthetic(args, torch, dist, rank, world, local, stream, hbm_peak, barrier, max_over_ranks)
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(n_sweeps=args.cpu_sweeps)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "fr1desk measurements from the committed fixture (TUM RGB-D / ORB-SLAM keyframes); synthetic BAL graph for the roofline",
            "config": {"workload": WORKLOAD, "step": "one 200-iteration solve from the initial state",
                       "msgs_per_step": msgs_per_step, "parallelism": "replicas only (fr1desk does not shard)" if world > 1 else "1 GPU",
                       "l2": "state reset + 512 MB L2 flush between timed steps; within a step the 10 MB state is L2-resident by nature",
                       "tile_edges": eng.tile_edges, "n_tiles": eng.n_tiles},
            "us_per_iteration": 1e3 * ms_per_step / N_ITERS,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_t / args.steps,
                    "what": "create_ba_graph from pinned host arrays (graph compile + upload) + priors + 200 synchronous iterations with the resets at 3 and 8 + final means on the host; wall clock"},
            "e2e_client_loop": {"value": loop_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_loop,
                                "ms_per_step": 1e3 * loop_t / args.steps, "max_rel_err": loop_parity,
                                "what": "the same solve driven like ba.py's loop body: 200 x (are, energy, relinearisation count and the viewer's means read back to the host, synchronous_iteration); a serial CPU<->GPU ping-pong"},
            "gpu_launches": launches,
            "parity": {"max_rel_err_means_vs_reference_fixture": parity_mu, "e2e_max_rel_err": e2e_parity, "tol": 1e-4,
                       "ok": bool(parity_mu < 1e-4 and e2e_parity < 1e-4 and loop_parity < 1e-4), "final_are_px": are_final},
            "roofline": roofline, "roofline_fr1desk": roof_fr1, "peak_source": peak_src,
            "cpu_baseline": cpu, "synthetic": synth,
        }
        emit(json.dumps(line))
    if dist is not None:
        # never let a teardown problem hold the box: the numbers are out, leave within 30 s
        import threading
        threading.Timer(30.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        os._exit(0)


def bench_synthetic(args, torch, dist, rank, world, local, stream, hbm_peak, barrier, max_over_ranks):
    """configs[3]/[4]: 1k keyframes / 1M landmarks / 10M factors; landmark-partitioned over the ranks."""
    from gbp_b200.synthetic import make_synthetic
    from gbp_b200.dist import PartitionedBAGraph
    t0 = time.perf_counter()
    prob = make_synthetic(args.synth_cams, args.synth_lmks, 10, seed=0)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    pg = PartitionedBAGraph(prob, CFG, rank=rank, world=world, device=local, stream=stream, dist=dist,
                            torch_stream=torch.cuda.current_stream(), p2p=args.p2p or None)
    build_s = time.perf_counter() - t0
    pg.generate_priors_var(CFG["prior_std_weaker_factor"])
    pg.update_all_beliefs()
    captured = pg.capture(local_relin=True, robustify=True) if not args.no_capture else False
    F, Lm, C = prob.n_edges, prob.n_points, prob.n_keyframes
    k = args.synth_iters
    for _ in range(max(args.warmup, 3)):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    barrier()
    l0 = pg.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    e1.record()
    torch.cuda.synchronize()
    launches = pg.engine.launch_count() - l0
    t = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    barrier()
    are, energy, nrel = pg.metrics()
    total_b, _ = b_alg(F, Lm, C)                                   # 696 B per factor
    eng = pg.engine
    _, sweep_b_survey = b_alg(eng.F, eng.L, eng.C)
    _, sweep_b_local = b_alg(eng.F, eng.L, eng.C, eng.msg_cam_width)   # what this engine's layout has to move
    total_b_layout, _ = b_alg(F, Lm, C, eng.msg_cam_width)
    tot_ms, sweep_ms = eng.time_iterations(k, True, True, per_kernel=True) if world == 1 else (None, None)
    # sustained: 200 back-to-back iterations (~0.3 s of continuous fp64 + HBM load; the burst above is ~30 ms)
    ks = args.synth_sustained
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(ks):
        pg.synchronous_iteration(robustify=True, local_relin=True)
    s1.record()
    torch.cuda.synchronize()
    t_sus = max_over_ranks(s0.elapsed_time(s1) / 1e3)
    barrier()
    synth = {"workload": f"synthetic BAL {C} keyframes / {Lm} landmarks / {F} factors, {'landmark-partitioned over %d GPUs, one exchange of keyframe partial sums per iteration' % world if world > 1 else '1 GPU'}",
             "value": k * 2 * F / t, "unit": UNIT, "ms_per_iteration": 1e3 * t / k, "iterations_timed": k, "scaling": "strong",
             "algorithmic_bytes_per_iteration": total_b, "achieved_gbs_whole_iteration_per_gpu": total_b / world / (t / k) / 1e9,
             "frac_of_hbm_peak_whole_iteration": total_b / world / (t / k) / 1e9 / hbm_peak,
             "sustained": {"iterations": ks, "ms_per_iteration": 1e3 * t_sus / ks, "value": ks * 2 * F / t_sus, "unit": UNIT,
                           "frac_of_hbm_peak_whole_iteration": total_b / world / (t_sus / ks) / 1e9 / hbm_peak,
                           "note": "back-to-back iterations for ~0.3 s; the burst figure above times %d iterations" % k},
             "l2": "%.1f GB streamed per iteration (inputs larger than L2, no flush needed)" % (total_b_layout / 1e9),
             "gpu_launches": launches, "are_px_after": are, "energy_after": energy, "generate_s": gen_s, "graph_build_s": build_s,
             "tile_edges": eng.tile_edges, "n_tiles_local": eng.n_tiles, "iteration_captured_in_cuda_graph": bool(captured) or world == 1,
             "exchange": None if world == 1 else ("peer-memory kernels (gbp_ba_p2p_*)" if pg.p2p else "NCCL all-gather"),
             "layout": {"msg_cam_doubles": eng.msg_cam_width, "sweep_kernel_build": eng.sweep_variant, "l2_prefetch_tiles": eng.prefetch_tiles,
                        "algorithmic_bytes_per_iteration_this_layout": total_b_layout,
                        "note": "factor->keyframe messages are stored with their rank-2 precision factored (18 doubles instead of 27): 144 B per factor less than SURVEY 8(d)'s 696 B; the SURVEY figure is kept for algorithmic_bytes_per_iteration / frac_of_hbm_peak_whole_iteration"}}
    roof = None
    if sweep_ms is not None:
        roof = roofline_entry(synth["workload"], sweep_ms / k, sweep_b_survey, sweep_b_local, hbm_peak,
                              ncu_traffic(f"sweep_kernel/synthetic_{C}_{Lm}_{F}"), eng.msg_cam_width)
        synth["ms_per_iteration_eager_with_events"] = tot_ms / k
    pg.close()
    return synth, roof


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


def cpu_baseline_sample(n_sweeps):
    """CPU baseline on the GPU box's host cores: the plain-C OpenMP port (oracle/gbp_oracle.c) on all threads, plus
    the single-threaded NumPy port for reference."""
    from oracle.gbp_oracle import BAOracle
    threads = os.cpu_count() or 1
    times, o, err = _c_oracle_solves(12, threads)
    times = times[2:]
    c_val = len(times) * N_ITERS * 2 * o.F / float(np.sum(times))
    prob, _ = load_fr1desk()
    n = BAOracle(prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG)
    n.generate_priors_var(CFG["prior_std_weaker_factor"])
    n.update_all_beliefs()
    n.synchronous_iteration(robustify=True, local_relin=True)
    t0 = time.perf_counter()
    for i in range(n_sweeps):
        n.synchronous_iteration(robustify=True, local_relin=True)
    dt = time.perf_counter() - t0
    return {"value": c_val, "unit": UNIT, "cores": o.threads, "kind": "port",
            "sample": f"{len(times)} complete 200-iteration fr1desk solves by oracle/gbp_oracle.c (plain C + OpenMP, reference arithmetic form), {float(np.sum(times)):.1f} s; means {err:.1e} from the reference fixture",
            "host_cpus": os.cpu_count(),
            "numpy_port_1_core": {"value": n_sweeps * 2 * n.F / dt, "unit": UNIT, "sample": f"{n_sweeps} iterations by oracle/gbp_oracle.py"},
            "reference_measured_in_build_container": {"value": 7780.0, "unit": UNIT, "cores": 1,
                                                      "source": "BASELINE.md: unmodified Python reference, 3.419 s per iteration on fr1desk"}}


def bench_reference(args):
    """Reference arm: the reference's algorithm on the host cores.  The Python reference cannot travel to the GPU box, so
    this times the plain-C OpenMP port of it (oracle/gbp_oracle.c, pinned against the reference's fixtures) with every
    host thread, one complete 200-iteration fr1desk solve per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times, o, err = _c_oracle_solves(args.warmup + args.steps, threads)
    times = times[args.warmup:]
    msgs = N_ITERS * 2 * o.F
    value = args.steps * msgs / float(np.sum(times))
    sample = f"one complete 200-iteration fr1desk solve per step by oracle/gbp_oracle.c (plain C + OpenMP port of the reference algorithm, {o.threads} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "fr1desk measurements from the committed fixture",
            "config": {"workload": WORKLOAD, "step": "one 200-iteration solve from the initial state", "msgs_per_step": msgs},
            "parity": {"max_rel_err_means_vs_reference_fixture": err},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": o.threads, "kind": "port", "sample": sample, "host_cpus": os.cpu_count()},
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
    ap.add_argument("--no-capture", action="store_true", help="multi-GPU: do not capture the iteration in a CUDA graph")
    ap.add_argument("--p2p", action="store_true", help="multi-GPU: peer-memory exchange kernels instead of the NCCL all-gather (experimental)")
    ap.add_argument("--synth-cams", type=int, default=1000)
    ap.add_argument("--synth-lmks", type=int, default=1_000_000)
    ap.add_argument("--synth-iters", type=int, default=20)
    ap.add_argument("--synth-sustained", type=int, default=200)
    ap.add_argument("--cpu-sweeps", type=int, default=40, help="NumPy-port iterations timed for the secondary CPU figure")
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
