#!/usr/bin/env python
"""Where does the resident kernel's state first differ from the two-kernel iteration?  (diagnostic, one GPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbp_b200 import _lib as L  # noqa: E402
from gbp_b200 import balio  # noqa: E402
from gbp_b200.ba import create_ba_graph  # noqa: E402

CFG = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
FIELDS = ("F_MSG_LMK", "F_MSG_CAM", "F_CAM_PARTIAL", "F_LMK_BELIEF", "F_CAM_BELIEF", "F_LINPOINT", "F_ITERS", "F_FLAGS")
name = sys.argv[1] if len(sys.argv) > 1 else "fr1desk"
G = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
prob = balio.BALProblem(G["in_cam_id"], G["in_lmk_id"], G["in_z"], G["in_cam0"], G["in_lmk0"], G["in_K"])
for n in (2, 3, 4, 9, 16, 17):
    gs = []
    for res in (0, 1):
        g = create_ba_graph(prob, CFG)
        g._eng.tune(L.TUNE_RESIDENT, res)
        g.generate_priors_var(50.0)
        g.update_all_beliefs()
        g._eng.fill_iters(8 if n >= 16 else 1)
        g.iterate(n, robustify=True, local_relin=True)
        gs.append(g)
    out = []
    for f in FIELDS:
        a, b = gs[0]._eng.read(getattr(L, f)).astype(float), gs[1]._eng.read(getattr(L, f)).astype(float)
        d = np.abs(a - b)
        rows = np.nonzero(d.max(axis=1) > 0)[0]
        out.append(f"{f[2:]}: max {d.max():.3e} rel {d.max() / max(np.abs(a).max(), 1e-300):.1e} rows {len(rows)}/{len(a)} cols {sorted(set(np.nonzero(d > 0)[1].tolist()))[:12]}")
    print(f"n={n}\n  " + "\n  ".join(out), flush=True)
    for g in gs:
        g.close()
