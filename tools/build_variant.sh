#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>  -> dots_ocr_b200/build/variants/lib_<name>.so (tuning A/B builds;
# select with DOTS_B200_LIB=<path>).  The in-tree product library is untouched.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=dots_ocr_b200/build/variants; mkdir -p $out/$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
for f in common gemm_tcgen05 attn_fwd_mma attn_fwd_tcgen05 attn_decode elementwise decode_gemm partition; do
  [ -f dots_ocr_b200/csrc/$f.cu ] && nvcc $FLAGS "$@" -c dots_ocr_b200/csrc/$f.cu -o $out/$name/$f.o &
done
wait
nvcc -shared -o $out/lib_$name.so $out/$name/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo $out/lib_$name.so
