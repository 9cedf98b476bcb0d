#!/bin/bash
# 2-GPU session: the NCCL partition test and the landmark-partitioned bench (driver-style torchrun launch).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== dist gpu test"; timeout 200 python -m pytest tests/test_dist_gpu.py -x -q --no-header -p no:cacheprovider > gpurun_out/dist_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/dist_tests.log
echo "== bench --gpus 2"; timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --synth-sustained 40 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2_stderr.log; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2_stderr.log
