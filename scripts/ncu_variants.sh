mkdir -p gpurun_out
for v in 6 7; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/ncu_sweep_v$v -f python scripts/profile_synth.py --variant $v --iters 8 > gpurun_out/ncu_v$v.log 2>&1; echo "v$v rc=$?"; tail -2 gpurun_out/ncu_v$v.log | head -1
done
