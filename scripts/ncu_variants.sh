#!/bin/bash
# ncu evidence for profiles/: one --set full capture of the default sweep and belief kernels on the 10 M-factor graph,
# and the launch list of the bench command (first 400 launches: the fr1desk part, 2 launches per iteration).
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|belief_kernel" -s 12 -c 2 -o gpurun_out/ncu_default_large -f python scripts/profile_synth.py --iters 8 > gpurun_out/ncu_default_large.log 2>&1; echo "full rc=$?"; grep "variant=" gpurun_out/ncu_default_large.log | cut -c1-300
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-synthetic > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_bench.csv
