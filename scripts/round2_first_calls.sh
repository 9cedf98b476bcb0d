#!/bin/bash
# First GPU calls of round 2: hardware runs of what round 1 left compiled-but-unrun.  Each block is one gpurun call.
#
#   1 GPU :  gpurun --timeout 300 -- 'bash scripts/round2_first_calls.sh one'
#   2 GPUs:  gpurun --gpus 2 --timeout 400 -- 'bash scripts/round2_first_calls.sh two'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
case "$1" in
one)
  echo "== one-kernel iteration (variant 11) + ring kernel: parity"; GBP_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_variants_gpu.py -q -x --no-header -p no:cacheprovider > gpurun_out/r2_exp_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_exp_tests.log
  echo "== fr1desk A/B: default vs variant 11"; timeout 60 python scripts/ab_variants.py --variants 11 --skip-synthetic --reps 9 > gpurun_out/r2_ab_v11.log 2>&1; echo "rc=$?"; cut -c1-400 gpurun_out/r2_ab_v11.log
  echo "== 10 M-factor graph: default (7 + prefetch) vs 12 (register column sums)"; timeout 120 python scripts/ab_variants.py --skip-fr1desk --variants 12 > gpurun_out/r2_ab_v12.log 2>&1; echo "rc=$?"; cut -c1-500 gpurun_out/r2_ab_v12.log
  echo "== end-to-end with / without the pooled arena"; for p in 0 1; do GBP_POOL_ALLOC=$p GBP_BENCH_DEBUG=1 timeout 120 python bench.py --no-synthetic --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_pool$p.json 2> gpurun_out/r2_bench_pool$p.err; echo "pool=$p rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_pool$p.json"))
print("value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "client loop ms", d["e2e_client_loop"]["ms_per_step"])
PY
  done
  ;;
two)
  echo "== 2-GPU partition test, NCCL and peer-memory exchange"; GBP_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_dist_gpu.py -q -x --no-header -p no:cacheprovider > gpurun_out/r2_dist_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_dist_tests.log
  for x in "" "--p2p"; do
    echo "== bench --gpus 2 $x"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --synth-sustained 40 --no-cpu-baseline $x > gpurun_out/r2_bench_n2$x.json 2> gpurun_out/r2_bench_n2$x.err; echo "rc=$?"
    python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n2$x.json") if l.startswith("{")][-1])
s = d["synthetic"]; print(s["exchange"], s["ms_per_iteration"], s["value"], s["are_px_after"])
PY
  done
  ;;
*) echo "usage: $0 one|two"; exit 2;;
esac
