#!/bin/bash
mkdir -p gpurun_out
bash tests/run_gpu.sh tests/test_attn_tc_gpu.py || exit 1
echo "== default"; timeout 300 python tools/bench_ops.py attn 2>&1 | grep '"tc"' | cut -c1-160
for v in "$@"; do
  echo "== $v"; DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_$v.so timeout 300 python -m pytest tests/test_attn_tc_gpu.py -x -q -m gpu 2>&1 | tail -1
  DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_$v.so timeout 300 python tools/bench_ops.py attn 2>&1 | grep '"tc"' | cut -c1-160
done
