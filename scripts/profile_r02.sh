#!/bin/bash
# Round-2 evidence, run on the GPU box (gpurun): every number first WITHOUT a profiler, ncu only afterwards.
# Outputs land in gpurun_out/; scripts/make_profile_summaries_r02.py turns them into the tracked files under profiles/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_n1.err
python scripts/config_bench.py all > gpurun_out/r02_config_bench.txt 2>&1
python scripts/small_probe.py > gpurun_out/r02_small_batch.txt 2>&1
PINN_FUSED_SMALL_ROUNDS=0 python scripts/small_probe.py > gpurun_out/r02_small_batch_off.txt 2>&1
python scripts/admm_cancellation_study.py > gpurun_out/r02_admm_cancellation.txt 2>&1
python scripts/tanh_variants.py > gpurun_out/r02_tanh_variants.txt 2>&1
PINN_B200_LIB=$PWD/scripts/variants/libpinn_trace.so python scripts/tc_phase_trace.py 128 75776 > gpurun_out/r02_tensor_phase_trace.txt 2>&1
PINN_B200_LIB=$PWD/scripts/variants/libpinn_sktrace.so python scripts/small_trace.py 1000 > gpurun_out/r02_small_kernel_trace.txt 2>&1
python -m pytest tests -m gpu -q -s --timeout 900 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
# ---- under ncu (never a bench value) ----
B="python bench.py --steps 2 --warmup 3 --nf-global 8388608 --nf-wide 262144 --cpu-points 65536"
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launch.log 2>&1
T="python scripts/tensor_bench.py 128 75776"
ncu --set full --clock-control none --import-source on -k regex:pinn_tc_kernel -s 2 -c 1 -o gpurun_out/prof_tensor_r02 -f $T > gpurun_out/r02_ncu_tensor.log 2>&1
Q="python scripts/quick_bench.py 2097152"
ncu --set full --clock-control none --import-source on -k regex:pinn_fused_kernel -s 3 -c 1 -o gpurun_out/prof_fused_r02 -f $Q > gpurun_out/r02_ncu_fused.log 2>&1
S="python scripts/quick_bench.py 1000"
ncu --set full --clock-control none --import-source on -k regex:pinn_fused_small_kernel -s 3 -c 1 -o gpurun_out/prof_small_r02 -f $S > gpurun_out/r02_ncu_small.log 2>&1
tail -2 gpurun_out/r02_ncu_*.log; tail -3 gpurun_out/r02_pytest_gpu.log; head -c 600 gpurun_out/r02_bench_n1.json
