#!/bin/bash
# One GPU-box session: A/B of the opt-in variants and their tests.  Every step has its own timeout and writes under
# gpurun_out/ so a cut-off call still leaves what finished.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== experimental tests"; GBP_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_experimental_gpu.py -q --no-header -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/exp_tests.log
echo "== ab_variants" ; timeout 300 python scripts/ab_variants.py $AB_ARGS > gpurun_out/ab_stdout.log 2> gpurun_out/ab_stderr.log; echo "rc=$?"; cut -c1-700 gpurun_out/ab_stdout.log; tail -5 gpurun_out/ab_stderr.log
