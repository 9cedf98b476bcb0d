for v in "" "NCCL_PROTO=LL" "NCCL_PROTO=LL128" "NCCL_PROTO=Simple" "NCCL_ALGO=Ring" "NCCL_NVLS_ENABLE=0"; do
  echo "== variant: [$v]"
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --no-parity-1gpu --no-fr1desk-replicas 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms/iter %.4f value %.4e phases %s' % (d['ms_per_iteration'], d['value'], {k:round(v,1) for k,v in d['iteration_phases_us'].items() if k!='note'}))
"
done
