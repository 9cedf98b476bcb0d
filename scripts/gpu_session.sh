#!/bin/bash
# One GPU-box session.  Every step has its own timeout and writes under gpurun_out/ so a cut-off call still leaves
# what finished.  STEPS selects: tests exp ab bench (default: all four).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
STEPS=${STEPS:-"tests exp ab bench"}
for s in $STEPS; do case $s in
tests) echo "== default gpu suite"; timeout 400 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/gpu_tests.log;;
exp) echo "== experimental tests"; GBP_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_variants_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/exp_tests.log;;
ab) echo "== ab_variants" ; timeout 300 python scripts/ab_variants.py $AB_ARGS > gpurun_out/ab_stdout.log 2> gpurun_out/ab_stderr.log; echo "rc=$?"; cut -c1-700 gpurun_out/ab_stdout.log; tail -5 gpurun_out/ab_stderr.log;;
bench) echo "== bench"; GBP_BENCH_DEBUG=1 timeout 280 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_stderr.log; echo "rc=$?"; cut -c1-3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_stderr.log;;
esac; done
