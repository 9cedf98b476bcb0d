#!/usr/bin/env python
"""A/B timings of the engine's tuning knobs on one GPU (a few seconds each):

  fr1desk   200-iteration solve (the bench's `value` workload), repeated runs
  synthetic 1k / 1M / 10M graph, per-kernel CUDA-event timing of the sweep: tile size, landmark block, L2 prefetch distance

Every run's final means are compared with the first run's (and fr1desk with the reference fixture).
One JSON object per line on stdout (and appended to gpurun_out/ab_variants.jsonl when that directory exists).

    python scripts/ab_variants.py --fr1desk --synthetic --tiles 32,64 --blocks 0,65536 --pf 600
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
from gbp_b200 import _lib as L  # noqa: E402
from gbp_b200 import balio  # noqa: E402
from gbp_b200.ba import create_ba_graph  # noqa: E402
from gbp_b200.synthetic import make_synthetic  # noqa: E402

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
OUT = os.path.join(ROOT, "gpurun_out", "ab_variants.jsonl")


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


def fr1desk(reps, variants=(0,)):
    G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
    mu_ref = np.concatenate([G["s199_cam_mu"], G["s199_lmk_mu"]])
    base = None
    runs = [(f"kv{v}_run{r}", v) for r in (1, 2) for v in variants]
    for tag, w in runs:
        try:
            g = create_ba_graph(prob, CFG, kernel_variant=w)
            e = g._eng
            times = []
            for it in range(reps + 2):
                g.reset()
                g.generate_priors_var(50.0)
                g.update_all_beliefs()
                e.synchronize()
                t0 = time.perf_counter()
                solve(e)
                e.synchronize()
                if it >= 2:
                    times.append(time.perf_counter() - t0)
            mu = g.get_means()
            if base is None:
                base = mu
            emit({"graph": "fr1desk", "run": tag, "ms_per_solve_wall_min": 1e3 * min(times), "ms_per_solve_wall_median": 1e3 * float(np.median(times)),
                  "us_per_iteration_min": 1e6 * min(times) / 200, "rel_err_vs_reference": float(np.max(np.abs(mu - mu_ref)) / np.max(np.abs(mu_ref))),
                  "max_abs_diff_vs_first": float(np.max(np.abs(mu - base))), "are": g.are()})
            g.close()
        except Exception as ex:      # noqa: BLE001
            emit({"graph": "fr1desk", "run": tag, "error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc()[-600:]})


def synthetic(cams, lmks, tiles, blocks, pfs, iters, lanes=(0,)):
    prob = make_synthetic(cams, lmks, 10, seed=0)
    base = None
    for T in tiles:
        for blk in blocks:
          for ln in lanes:
            for pf in pfs:
                tag = f"T{T}_blk{blk}_pf{pf}_lanes{ln}"
                try:
                    g = create_ba_graph(prob, CFG, tile_edges=T, lmk_block=blk)
                    e = g._eng
                    if ln:
                        e.tune(L.TUNE_BELIEF_LANES, ln)
                    if pf >= 0:
                        e.tune(L.TUNE_PREFETCH_TILES, pf)
                    g.generate_priors_var(50.0)
                    g.update_all_beliefs()
                    e.iterate(5, True, True)
                    tot, sw = e.time_iterations(iters, True, True, per_kernel=True)
                    tot_g, _ = e.time_iterations(iters, True, True, per_kernel=False)
                    mu = g.get_means()
                    if base is None:
                        base = mu
                    F = e.F
                    emit({"graph": f"synthetic {cams}/{lmks}/{F}", "run": tag, "tiles": e.n_tiles, "tile_edges": e.tile_edges, "slots": e.n_slots,
                          "padding_frac": e.n_slots / F - 1.0, "prefetch_tiles": e.prefetch_tiles, "sweep_ms": sw / iters,
                          "iteration_ms_events": tot / iters, "iteration_ms_graph": tot_g / iters, "belief_ms": (tot - sw) / iters,
                          "msgs_per_s_graph": 2 * F / (tot_g / iters * 1e-3),
                          "are": g.are(), "max_rel_diff_means_vs_first": float(np.max(np.abs(mu - base)) / np.max(np.abs(base)))})
                    g.close()
                except Exception as ex:      # noqa: BLE001
                    emit({"graph": "synthetic", "run": tag, "error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc()[-600:]})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fr1desk", action="store_true")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--cams", type=int, default=1000)
    ap.add_argument("--lmks", type=int, default=1_000_000)
    ap.add_argument("--tiles", default="0")
    ap.add_argument("--blocks", default="0")
    ap.add_argument("--pf", default="-1", help="L2 prefetch distances in tiles (-1 = automatic)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--lanes", default="0", help="lanes per landmark of the belief kernel (1, 8, 32; 0 = automatic)")
    ap.add_argument("--kv", default="0", help="fr1desk: kernel_variant values to run (0 automatic, 1 full rows, 2 streaming build)")
    ap.add_argument("--lib", default="", help="load this build of libgbp_b200.so instead of the in-tree one (A/B of two builds on one box)")
    a = ap.parse_args()
    if a.lib:
        L.LIB_PATH = os.path.abspath(a.lib)
        emit({"lib": L.LIB_PATH})
    if a.fr1desk:
        fr1desk(a.reps, [int(x) for x in a.kv.split(",")])
    if a.synthetic:
        synthetic(a.cams, a.lmks, [int(x) for x in a.tiles.split(",")], [int(x) for x in a.blocks.split(",")],
                  [int(x) for x in a.pf.split(",")], a.iters, [int(x) for x in a.lanes.split(",")])


if __name__ == "__main__":
    main()
