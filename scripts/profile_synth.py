#!/usr/bin/env python
"""Small driver for ncu: build a synthetic BAL graph and run a few synchronous iterations.

    ncu --set full -k regex:sweep_kernel -s 6 -c 2 ... python scripts/profile_synth.py --lmks 1000000 --iters 8
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gbp_b200.ba import create_ba_graph  # noqa: E402
from gbp_b200.synthetic import make_synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cams", type=int, default=1000)
ap.add_argument("--lmks", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--fr1desk", action="store_true")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--loss", default=None)
a = ap.parse_args()
cfg = dict(gauss_noise_std=2, loss=a.loss, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
if a.fr1desk:
    import numpy as np
    from gbp_b200 import balio
    G = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fr1desk.npz"))
    prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
else:
    prob = make_synthetic(a.cams, a.lmks, 10, seed=0)
g = create_ba_graph(prob, cfg, tile_edges=a.tile, lmk_block=a.block, kernel_variant=a.variant)
g.generate_priors_var(50.0)
g.update_all_beliefs()
e = g._eng
for i in range(a.iters):                      # eager launches (one sweep_kernel + one belief_kernel each)
    e.sweep_local(1 | 2 | 4 | 8 | 16)
    e.cam_update()
e.synchronize()
tot, sw = e.time_iterations(a.iters, True, True, per_kernel=True)
tot_g, _ = e.time_iterations(a.iters, True, True, per_kernel=False)
import time as _t
e.synchronize(); _t0 = _t.perf_counter(); e.iterate(200, True, True); e.synchronize(); it200 = (_t.perf_counter() - _t0) / 200 * 1e3
F, L, C = e.F, e.L, e.C
print(f"variant={a.variant} F={F} L={L} C={C} tiles={e.n_tiles}x{e.tile_edges}  sweep {sw / a.iters:.4f} ms  iteration eager {tot / a.iters:.4f} ms  graph {tot_g / a.iters:.4f} ms  iterate(200) wall {it200:.4f} ms/iter"
      f"  sweep GB/s {(696 * F + 96 * L + 264 * C) / (sw / a.iters * 1e-3) / 1e9:.1f}  iter GB/s {(696 * F + 264 * L + 744 * C) / (tot_g / a.iters * 1e-3) / 1e9:.1f}"
      f"  ARE {g.are():.4f}")
