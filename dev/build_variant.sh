#!/bin/bash
# usage: bash dev/build_variant.sh <name> <extra nvcc flags...>  -> dev/variants/lib_<name>.so (run with CONCEPTHASH_B200_LIB=...)
set -e
N="$1"; shift
mkdir -p dev/variants
S=concepthash_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default \
  --shared -cudart shared -I include "$@" -o dev/variants/lib_$N.so $S/api.cu $S/pack.cu $S/hist.cu $S/select_tc.cu $S/cand.cu $S/select_ap.cu $S/host_pack.cpp
echo built dev/variants/lib_$N.so
