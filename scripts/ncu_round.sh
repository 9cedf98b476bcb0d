#!/bin/bash
# ncu captures of one round (one GPU; never under a multi-rank launch).  TAG names the outputs under gpurun_out/:
#   ${TAG}_ncu_sweep_large.ncu-rep / _belief_large  ncu --set full of the dominant kernel and the belief kernel on the 10 M-factor graph
#   ${TAG}_ncu_fr1desk.ncu-rep                       both kernels on the headline graph (L2-resident)
#   ${TAG}_launches_bench.csv                        launch list (gpu__time_duration.sum) of the bench command
# Read the reports in the build container: python scripts/ncu_summary.py gpurun_out/X.ncu-rep --md profiles/X.md
mkdir -p gpurun_out
TAG=${TAG:-r2}
export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none --import-source on"
echo "== ncu sweep_kernel, 10 M-factor graph"
timeout 900 $NCU -k regex:sweep_kernel -s 5 -c 1 -f -o gpurun_out/${TAG}_ncu_sweep_large python scripts/profile_synth.py --iters 6 > gpurun_out/${TAG}_ncu_sweep_large.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_sweep_large.log
echo "== ncu belief_kernel, 10 M-factor graph"
timeout 900 $NCU -k regex:belief_kernel -s 5 -c 1 -f -o gpurun_out/${TAG}_ncu_belief_large python scripts/profile_synth.py --iters 6 > gpurun_out/${TAG}_ncu_belief_large.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_belief_large.log
echo "== ncu fr1desk (sweep + belief)"
timeout 600 $NCU -k regex:"sweep_kernel|belief_kernel" -s 10 -c 2 -f -o gpurun_out/${TAG}_ncu_fr1desk python scripts/profile_synth.py --fr1desk --iters 8 > gpurun_out/${TAG}_ncu_fr1desk.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_fr1desk.log
echo "== launch list of the bench command"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba-py --synth-iters 5 --synth-sustained 10 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_bench.csv > gpurun_out/${TAG}_launches_bench_summary.md; tail -30 gpurun_out/${TAG}_launches_bench_summary.md
