#!/usr/bin/env python
"""Where the end-to-end time of the fr1desk client flow goes (host side)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import balio
from gbp_b200.ba import create_ba_graph
import cProfile, pstats

G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk.npz"))
prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
STREAM = None
if "--torch-stream" in sys.argv:
    import torch
    _ws = torch.cuda.Stream(); torch.cuda.set_stream(_ws); STREAM = _ws.cuda_stream
    _fb = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
cfg = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)

def once(profile=False):
    t = [time.perf_counter()]
    g = create_ba_graph(prob, cfg, stream=STREAM); t.append(time.perf_counter())
    g.generate_priors_var(50.0); g.update_all_beliefs(); g._eng.synchronize(); t.append(time.perf_counter())
    tm = ts = tv = 0.0
    for i in range(200):
        if i in (3, 8):
            g._flush(); g._eng.fill_iters(1); g._invalidate((7,))
        a = time.perf_counter(); g.metrics(); b = time.perf_counter()
        g.cam_nodes[0].mu; g.lmk_nodes[0].mu; c = time.perf_counter()
        g.synchronous_iteration(robustify=True, local_relin=True); d = time.perf_counter()
        tm += b - a; tv += c - b; ts += d - c
    g._eng.synchronize(); t.append(time.perf_counter())
    m = g.get_means(); t.append(time.perf_counter())
    g.close(); t.append(time.perf_counter())
    return np.diff(t) * 1e3, tm * 1e3, tv * 1e3, ts * 1e3

for k in range(4):
    if STREAM is not None:
        _fb.zero_(); torch.cuda.synchronize()
    d, tm, tv, ts = once()
    print(f"create {d[0]:.2f} ms  priors+beliefs {d[1]:.2f}  loop {d[2]:.2f} (metrics {tm:.2f}, mu reads {tv:.2f}, sync_iter {ts:.2f})  final means {d[3]:.2f}  close {d[4]:.2f}  total {d.sum():.2f}")
pr = cProfile.Profile(); pr.enable(); once(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
