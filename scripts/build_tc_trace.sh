#!/bin/bash
# builds scripts/variants/libpinn_trace.so: the library with the tcgen05 kernel's phase tracing (-DPINN_TC_TRACE) for scripts/tc_phase_trace.py
set -e
cd "$(dirname "$0")/../pinns_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../scripts/variants /tmp/pinn_variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPINN_TC_TRACE -c pinn_tensor.cu -o /tmp/pinn_variants/tensor_trace.o 2>/dev/null
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/variants/libpinn_trace.so pinn_capi.o pinn_generic.o pinn_aux.o pinn_fused.o /tmp/pinn_variants/tensor_trace.o
echo built libpinn_trace.so
# scripts/variants/libpinn_sktrace.so: the small-batch fused kernel's phase tracing (-DPINN_FUSED_SMALL_TRACE) for scripts/small_trace.py
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPINN_FUSED_SMALL_TRACE -c pinn_fused.cu -o /tmp/pinn_variants/fused_sktrace.o 2>/dev/null
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/variants/libpinn_sktrace.so pinn_capi.o pinn_generic.o pinn_aux.o /tmp/pinn_variants/fused_sktrace.o pinn_tensor.o
echo built libpinn_sktrace.so
