#!/usr/bin/env python
"""Why does every graph after the first one of a process iterate ~5 % slower on fr1desk (8.9 vs 9.35 us per iteration)?
Times the 200-iteration solve (device events) for a sequence of graphs under different conditions."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import _lib as L  # noqa: E402
from gbp_b200 import balio  # noqa: E402
from gbp_b200.engine import BAEngine  # noqa: E402

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
    ev, _ = e.time_iterations(192, True, True, per_kernel=False)
    print(f"{tag:58s} wall min {1e6 * min(ts) / 200:.3f} us/iter   events(192 x 1-iteration graphs) {1e3 * ev / 192:.3f} us/iter", flush=True)


lib = L.load()
a = BAEngine(*args); timed(a, "A: first graph of the process")
b = BAEngine(*args); timed(b, "B: second graph, A still alive (own cudaMalloc)")
timed(a, "A again (B alive)")
a.close(); b.close()
c = BAEngine(*args); timed(c, "C: after closing A and B (arena + graphs from the shell)")
c.close()
lib.gbp_cache_configure(0, 0)
d = BAEngine(*args); timed(d, "D: shell cache off (fresh cudaMalloc, fresh graphs)")
d.close()
import torch
s = torch.cuda.Stream()
e = BAEngine(*args, stream=s.cuda_stream); timed(e, "E: cache off, on a torch stream")
e.close()
