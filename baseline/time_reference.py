"""Time the UNMODIFIED reference (joeaortiz/gbp) on this host: run with cwd = baseline/_ref (the staged copy of the reference
tree, see __graft_entry__.stage_reference), so that `from gbp import gbp_ba` is the reference's own package.

    cd baseline/_ref && python ../time_reference.py data/fr1desk.txt 5

Replicates ba.py:51-76 + the sweep call of ba.py:105 (the viewer lines are the only thing left out: trimesh / pyglet are
not installed), times `synchronous_iteration(robustify=True, local_relin=True)` (gbp/gbp.py:86-92) with perf_counter and
prints one JSON line.  Single-threaded by construction (pure Python + 9x9 NumPy blocks; BLAS threads pinned to 1)."""
import json
import os
import sys
import time

for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[v] = "1"
sys.path.insert(0, os.getcwd())
import warnings

warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402
from gbp import gbp_ba  # noqa: E402  (the reference's, from cwd)

bal_file = sys.argv[1] if len(sys.argv) > 1 else "data/fr1desk.txt"
n_sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
configs = dict(gauss_noise_std=2, loss=None, Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4,
               prior_std_weaker_factor=50.0)
t0 = time.perf_counter()
graph = gbp_ba.create_ba_graph(bal_file, configs)
t_create = time.perf_counter() - t0
graph.generate_priors_var(weaker_factor=50.0)
graph.update_all_beliefs()
times = []
for i in range(n_sweeps):
    t0 = time.perf_counter()
    graph.synchronous_iteration(robustify=True, local_relin=True)
    times.append(time.perf_counter() - t0)
F = len(graph.factors)
print(json.dumps({"n_factors": F, "n_keyframes": len(graph.cam_nodes), "n_landmarks": len(graph.lmk_nodes), "create_s": t_create,
                  "iteration_s": times, "median_iteration_s": float(np.median(times)), "msgs_per_s": 2 * F / float(np.median(times)),
                  "are_after": float(graph.are()), "module": os.path.abspath(gbp_ba.__file__), "numpy": np.__version__}))
