#!/bin/bash
# builds libpinn_b200 with each tanh variant of the fused kernel into scripts/variants/ (A/B on the GPU: scripts/tanh_variants.py)
set -e
cd "$(dirname "$0")/../pinns_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../scripts/variants /tmp/pinn_variants
for v in "t0_tanhf:-DPINN_FUSED_TANH=0" "t1_round1:-DPINN_FUSED_TANH=1" "t2_60_default:-DPINN_FUSED_TANH=2 -DPINN_FUSED_TANH_T=60" \
         "t2_60_no_newton:-DPINN_FUSED_TANH=2 -DPINN_FUSED_TANH_T=60 -DPINN_FUSED_TANH_NO_NEWTON" \
         "t2_55_own_fit4:-DPINN_FUSED_TANH=2 -DPINN_FUSED_TANH_T=55" "t2_56_own_fit5:-DPINN_FUSED_TANH=2 -DPINN_FUSED_TANH_T=56"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c pinn_fused.cu -o /tmp/pinn_variants/fused_$name.o 2>/dev/null
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/variants/libpinn_$name.so pinn_capi.o pinn_generic.o pinn_aux.o /tmp/pinn_variants/fused_$name.o pinn_tensor.o
  echo built $name
done
