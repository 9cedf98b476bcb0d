#!/usr/bin/env python
"""Follow-up of diag_second_graph.py: which process-global event makes fr1desk iterate ~5 % slower -- a second stream, a second
device allocation, or a second graph?  One scenario per process: python diag_second_graph2.py stream|malloc|graph_same_stream|graph_own_stream"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import balio  # noqa: E402
from gbp_b200.engine import BAEngine  # noqa: E402
import torch  # noqa: E402

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
G = np.load(os.path.join(ROOT, "tests", "golden", "fr1desk.npz"))
prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
args = (prob.cam_id, prob.lmk_id, prob.z, prob.cam_means, prob.lmk_means, prob.K4, CFG)


def timed(e, tag, reps=12):
    ts = []
    for it in range(reps + 2):
        e.reset(); e.generate_priors(50.0); e.update_beliefs(); e.synchronize()
        t0 = time.perf_counter()
        e.iterate(3, True, True); e.fill_iters(1); e.iterate(5, True, True); e.fill_iters(1); e.iterate(192, True, True)
        e.synchronize()
        if it >= 2:
            ts.append(time.perf_counter() - t0)
    print(f"{tag:70s} {1e6 * min(ts) / 200:.3f} us/iter", flush=True)


what = sys.argv[1]
torch.cuda.init()
s0 = torch.cuda.Stream()
a = BAEngine(*args, stream=s0.cuda_stream)
timed(a, f"[{what}] A on the only non-default stream")
if what == "stream":
    s1 = torch.cuda.Stream()
    with torch.cuda.stream(s1):
        x = torch.zeros(16, device="cuda"); x += 1
    torch.cuda.synchronize()
    timed(a, f"[{what}] A after a second stream ran one kernel")
elif what == "malloc":
    x = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(s0):
        x.fill_(1)
    torch.cuda.synchronize()
    timed(a, f"[{what}] A after a 32 MB allocation + fill on the same stream")
elif what == "graph_same_stream":
    b = BAEngine(*args, stream=s0.cuda_stream)
    timed(b, f"[{what}] B (second graph, same stream)")
    timed(a, f"[{what}] A after B")
elif what == "graph_own_stream":
    b = BAEngine(*args)
    timed(b, f"[{what}] B (second graph, its own stream)")
    timed(a, f"[{what}] A after B")
elif what == "flush":
    x = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(s0):
        x.fill_(1)
    torch.cuda.synchronize()
    timed(a, f"[{what}] A after a 512 MB fill (L2 flush) on the same stream")
