#!/bin/bash
# usage: build_variant.sh <name> [-DPBN_...=...]...   -> ../variants/libpbn_<name>.so  (tuning experiments)
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p build_$name ../variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $@"
for f in runtime pair_f64 pair_f32 pair_gskip_f64 pair_gskip_f32 pair_shift_f64 pair_shift_f32 ucv_kernel ucv_host cv lg segmented cdf_sample spatial; do /usr/local/cuda/bin/nvcc $FLAGS -c $f.cu -o build_$name/$f.o 2>build_$name.$f.log & done
/usr/local/cuda/bin/nvcc $FLAGS -c host_util.cpp -o build_$name/host_util.o & wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libpbn_$name.so build_$name/*.o -lcudart
rm -rf build_$name
