# launch list of the bench command + one full ncu capture of the fused kernel (each only after the plain run exited 0)
set -x
B="python bench.py --steps 3 --warmup 3 --nf 4194304 --cpu-points 65536"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
Q="python scripts/quick_bench.py 2097152"
$Q > gpurun_out/plain_quick.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pinn_fused_kernel -s 3 -c 1 -o gpurun_out/prof_fused_r01d -f $Q > gpurun_out/ncu_fused.log 2>&1
cat gpurun_out/plain_quick.log
tail -2 gpurun_out/ncu_launch.log gpurun_out/ncu_fused.log
