#!/bin/bash
# timing-only A/B of library variants on the attention micro-bench
echo "== default"; timeout 300 python tools/bench_ops.py attn 2>&1 | grep '"tc"' | cut -c1-160
for v in "$@"; do
  echo "== $v"; DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_$v.so timeout 300 python tools/bench_ops.py attn 2>&1 | grep '"tc"' | cut -c1-160
done
