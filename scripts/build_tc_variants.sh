#!/bin/bash
# A/B builds of the tcgen05 kernel (measurement knobs) into scripts/variants/
set -e
cd "$(dirname "$0")/../pinns_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../scripts/variants /tmp/pinn_variants
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c pinn_tensor.cu -o /tmp/pinn_variants/tensor_$name.o 2>/dev/null
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/variants/libpinn_tc_$name.so pinn_capi.o pinn_generic.o pinn_aux.o pinn_fused.o /tmp/pinn_variants/tensor_$name.o
  echo built $name
done
