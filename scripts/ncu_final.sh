mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=r3c
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:sweep_kernel -s 5 -c 1 -f -o gpurun_out/${TAG}_ncu_sweep_large python scripts/profile_synth.py --iters 6 > gpurun_out/${TAG}_ncu_sweep_large.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_sweep_large.log
timeout 600 $NCU -k regex:"sweep_kernel|belief_kernel" -s 10 -c 2 -f -o gpurun_out/${TAG}_ncu_fr1desk python scripts/profile_synth.py --fr1desk --iters 8 > gpurun_out/${TAG}_ncu_fr1desk.log 2>&1; echo "rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba-py --synth-iters 5 --synth-sustained 10 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_bench.csv > gpurun_out/${TAG}_launches_bench_summary.md; head -12 gpurun_out/${TAG}_launches_bench_summary.md
