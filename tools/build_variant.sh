#!/bin/bash
# builds rf_inv_b200/librfinv_b200_<name>.so with extra nvcc flags (experiments): tools/build_variant.sh name -DFLAG ...
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr "$@" \
  -shared -o rf_inv_b200/librfinv_b200_$name.so rf_inv_b200/csrc/{capi,forward,forward_general,likelihood,pt,comm,fp64_peak,host_io}.cu -lcudart -ldl 2>&1 | grep -E "error" 
ls -la rf_inv_b200/librfinv_b200_$name.so
