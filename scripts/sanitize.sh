#!/bin/bash
# compute-sanitizer runs of the product kernels: memcheck, racecheck and synccheck of smoke() (fr1desk_vsmall:
# the L2-resident build of the sweep kernel, belief kernel with 32 lanes, metrics, priors) and of a few iterations of a graph
# large enough for the streaming build (bulk copies of whole tiles, factored messages, L2 prefetch, chunked keyframe sums,
# belief kernel with 1 lane) and of the linear-factor path.  Logs -> gpurun_out/${TAG}_sanitize_*.log
mkdir -p gpurun_out
TAG=${TAG:-r2}
export PYTHONUNBUFFERED=1
SMOKE='import __graft_entry__ as g; g.smoke()'
LARGE='
import numpy as np
from gbp_b200.synthetic import make_synthetic
from gbp_b200.ba import create_ba_graph
cfg = dict(gauss_noise_std=2, loss="huber", Nstds=3.0, beta=0.01, num_undamped_iters=6, min_linear_iters=8, eta_damping=0.4)
prob = make_synthetic(40, 70000, 9, seed=1)          # 630 k factors -> 9.9 k tiles of 64: streaming build, 8 landmark chunks
g = create_ba_graph(prob, cfg, chunks=(8, 0, 8, 0, prob.n_points))
e = g._eng
print("tiles", e.n_tiles, "x", e.tile_edges, "build", e.sweep_variant, "chunks", e.lmk_chunks, "prefetch", e.prefetch_tiles)
assert e.sweep_variant == 2 and e.lmk_chunks == 8
g.generate_priors_var(50.0); g.update_all_beliefs()
for i in range(3):
    g.synchronous_iteration(robustify=True, local_relin=True)
print("ARE", g.are(), "energy", g.energy())
g.close()
'
for tool in memcheck racecheck synccheck; do
  for what in smoke large; do
    case $what in smoke) CODE="$SMOKE";; large) CODE="$LARGE";; esac
    echo "== compute-sanitizer --tool $tool : $what"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "$CODE" > gpurun_out/${TAG}_sanitize_${tool}_${what}.log 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|smoke:|^ARE|tiles" gpurun_out/${TAG}_sanitize_${tool}_${what}.log | tail -5
  done
done
echo "== compute-sanitizer --tool memcheck : linear-factor path (unmodified ndim_posegraph.py)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m gbp_b200.run baseline/_ref/ndim_posegraph.py --n_varnodes 30 --dim 3 > gpurun_out/${TAG}_sanitize_memcheck_lin.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/${TAG}_sanitize_memcheck_lin.log | tail -2
