#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]  ->  build/variants/NAME.so  (A/B tuning builds of the same library; load with OC_B200_LIB)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I include -Xcompiler -fPIC"
nvcc $F "$@" -c jaxovercooked_b200/csrc/oc_b200.cu -o $out/oc_b200.o &
if [ ! -f build/oc_policy.o ] || [ jaxovercooked_b200/csrc/oc_policy_kernels.cuh -nt build/oc_policy.o ] || [ jaxovercooked_b200/csrc/oc_policy.cu -nt build/oc_policy.o ]; then
  nvcc $F -c jaxovercooked_b200/csrc/oc_policy.cu -o build/oc_policy.o &
fi
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $out/oc_b200.o build/oc_policy.o
echo built build/variants/$name.so
