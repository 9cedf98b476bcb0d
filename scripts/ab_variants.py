#!/usr/bin/env python
"""A/B timings of the opt-in engine variants against the default engine (one GPU, a few seconds each):

  fr1desk   200-iteration solve (the bench's `value` workload), default vs GBP_PDL=1 vs kernel_variant 5
  synthetic 1k / 1M / 10M graph, per-kernel CUDA-event timing of the sweep, kernel_variant 0 vs 5

Every variant's final means are compared with the default engine's (and fr1desk with the reference fixture).
One JSON object per line on stdout (and appended to gpurun_out/ab_variants.jsonl when that directory exists).
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import balio  # noqa: E402
from gbp_b200.ba import create_ba_graph  # noqa: E402
from gbp_b200.synthetic import make_synthetic  # noqa: E402

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
OUT = os.path.join(ROOT, "gpurun_out", "ab_variants.jsonl")
VARIANTS, PDL, PF, RUNS = [5, 6, 7, 9], False, [0], []


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    if os.path.isdir(os.path.dirname(OUT)):
        with open(OUT, "a") as f:
            f.write(line + "\n")


def solve(e):
    e.iterate(3, True, True); e.fill_iters(1)
    e.iterate(5, True, True); e.fill_iters(1)
    e.iterate(192, True, True)


def fr1desk(reps):
    G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
    mu_ref = np.concatenate([G["s199_cam_mu"], G["s199_lmk_mu"]])
    base = None
    runs = [("default", "0", 0)] + [(f"variant{v}", "0", v) for v in VARIANTS] + [("default_again", "0", 0)]
    if PDL:
        runs.insert(1, ("pdl", "1", 0))
    for name, env, variant in runs:
        try:
            os.environ["GBP_PDL"] = env
            g = create_ba_graph(prob, CFG, kernel_variant=variant)
            e = g._eng
            ms = []
            for r in range(reps + 2):
                g.reset(); g.generate_priors_var(50.0); g.update_all_beliefs(); e.synchronize()
                # the engine runs on its own stream: host-side synchronisation on both sides + wall clock
                e.synchronize(); w0 = time.perf_counter()
                solve(e)
                e.synchronize(); w1 = time.perf_counter()
                if r >= 2:
                    ms.append(1e3 * (w1 - w0))
            mu = g.get_means()
            if base is None:
                base = mu
            emit({"graph": "fr1desk", "variant": name, "ms_per_solve_wall_min": min(ms), "ms_per_solve_wall_median": float(np.median(ms)),
                  "us_per_iteration_min": 1e3 * min(ms) / 200, "rel_err_vs_reference": float(np.max(np.abs(mu - mu_ref)) / np.max(np.abs(mu_ref))),
                  "max_abs_diff_vs_default": float(np.max(np.abs(mu - base))), "are": g.are()})
            g.close()
        except Exception as ex:                                  # keep going: the other variants are still worth timing
            emit({"graph": "fr1desk", "variant": name, "error": repr(ex), "trace": traceback.format_exc()[-600:]})
    os.environ["GBP_PDL"] = "0"


def synthetic(cams, lmks, iters):
    prob = make_synthetic(cams, lmks, 10, seed=0)
    base = None
    runs = [("default", 0, 0)] + [(f"variant{v}" + (f"_pf{d}" if d else ""), v, d) for v in VARIANTS for d in PF] + [("default_again", 0, 0)]
    if RUNS:
        runs = [("default", 0, 0)] + [(f"variant{v}_pf{d}", v, d) for v, d in RUNS] + [("default_again", 0, 0)]
    for name, variant, pf in runs:
        try:
            os.environ["GBP_PF_DIST"] = str(pf)
            g = create_ba_graph(prob, CFG, kernel_variant=variant)
            e = g._eng
            g.generate_priors_var(50.0); g.update_all_beliefs()
            e.iterate(4, True, True); e.synchronize()
            tot, sw = e.time_iterations(iters, True, True, per_kernel=True)
            tot_g, _ = e.time_iterations(iters, True, True, per_kernel=False)
            F, L, C = e.F, e.L, e.C
            per_edge = 540 if variant in (5, 7, 8) else 684          # bytes the sweep really moves per edge (ids 4, z 16, linpoint 72, iters/flags 16, messages)
            mu = g.get_means()
            if base is None:
                base = mu
            emit({"graph": f"synthetic {C}/{L}/{F}", "variant": name, "tiles": e.n_tiles, "tile_edges": e.tile_edges,
                  "sweep_ms": sw / iters, "iteration_ms_events": tot / iters, "iteration_ms_graph": tot_g / iters,
                  "sweep_alg_gbs": (696 * F + 96 * L + 264 * C) / (sw / iters * 1e-3) / 1e9,
                  "sweep_moved_gbs_estimate": (per_edge * F + 96 * L + 264 * C) / (sw / iters * 1e-3) / 1e9,
                  "msgs_per_s_graph": 2 * F / (tot_g / iters * 1e-3), "are": g.are(),
                  "max_rel_diff_means_vs_default": float(np.max(np.abs(mu - base)) / np.max(np.abs(base)))})
            g.close()
        except Exception as ex:
            emit({"graph": "synthetic", "variant": name, "error": repr(ex), "trace": traceback.format_exc()[-600:]})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--cams", type=int, default=1000)
    ap.add_argument("--lmks", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--skip-synthetic", action="store_true")
    ap.add_argument("--variants", default="5,6,7,9")
    ap.add_argument("--pdl", action="store_true")
    ap.add_argument("--pf", default="0", help="L2 prefetch distances (tiles) to try on the synthetic graph")
    ap.add_argument("--skip-fr1desk", action="store_true")
    ap.add_argument("--runs", default="", help="explicit synthetic runs: variant:pf,variant:pf,...")
    a = ap.parse_args()
    VARIANTS = [int(v) for v in a.variants.split(",") if v]
    PDL = a.pdl
    PF = [int(v) for v in a.pf.split(",") if v]
    RUNS = [tuple(int(x) for x in r.split(":")) for r in a.runs.split(",") if r]
    if not a.skip_fr1desk:
        fr1desk(a.reps)
    if not a.skip_synthetic:
        synthetic(a.cams, a.lmks, a.iters)
