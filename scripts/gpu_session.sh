#!/bin/bash
# One GPU-box session.  Every step has its own timeout and writes under gpurun_out/ so a cut-off call still leaves
# what finished.  STEPS selects: tests bench bench2 tests2 (default: tests bench).  TAG names the output files.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
STEPS=${STEPS:-"tests bench"}
TAG=${TAG:-r2}
for s in $STEPS; do case $s in
tests) echo "== default gpu suite"; timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${TAG}_gpu_tests.log;;
bench) echo "== bench"; GBP_BENCH_DEBUG=1 timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "rc=$?"; cut -c1-6000 gpurun_out/${TAG}_bench_n1.json; tail -45 gpurun_out/${TAG}_bench_n1.err;;
refarm) echo "== bench --impl reference"; timeout 600 python bench.py --impl reference ${BENCH_ARGS} > gpurun_out/${TAG}_bench_ref_n1.json 2> gpurun_out/${TAG}_bench_ref_n1.err; echo "rc=$?"; cut -c1-2000 gpurun_out/${TAG}_bench_ref_n1.json;;
tests2) echo "== 2-GPU partition tests (native NCCL exchange, C++ shard client)"; timeout 400 python -m pytest tests/test_dist_gpu.py tests/test_shard_client_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/${TAG}_dist_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${TAG}_dist_tests.log;;
benchN) N=${NGPU:-2}; for x in ""; do
    echo "== bench --gpus $N $x"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N ${BENCH_ARGS} $x > gpurun_out/${TAG}_bench_n${N}${x}.json 2> gpurun_out/${TAG}_bench_n${N}${x}.err; echo "rc=$?"
    cut -c1-5000 gpurun_out/${TAG}_bench_n${N}${x}.json; tail -8 gpurun_out/${TAG}_bench_n${N}${x}.err
  done;;
esac; done
