#!/bin/bash
# One GPU-box session: A/B of the opt-in variants, their tests, the bench line, then the default GPU suite.
# Every step has its own timeout and writes under gpurun_out/ so a cut-off call still leaves what finished.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== ab_variants" ; timeout 200 python scripts/ab_variants.py > gpurun_out/ab_stdout.log 2> gpurun_out/ab_stderr.log; echo "rc=$?"; cat gpurun_out/ab_stdout.log | cut -c1-600
echo "== experimental tests"; GBP_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_experimental_gpu.py -q -x --no-header -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/exp_tests.log
echo "== bench"; timeout 280 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_stderr.log; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_stderr.log
echo "== default gpu suite"; timeout 400 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/gpu_tests.log
echo "== ncu factored sweep (bonus)"; timeout 170 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/ncu_sweep_factored -f python scripts/profile_synth.py --variant 5 --iters 8 > gpurun_out/ncu_factored.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_factored.log
