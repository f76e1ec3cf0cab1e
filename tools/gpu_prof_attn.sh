#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05 -s 3 -c 1 -o gpurun_out/prof_attn_tc_$1 -f \
    python tools/bench_ops.py attn > gpurun_out/ncu_attn_tc_$1.log 2>&1; tail -3 gpurun_out/ncu_attn_tc_$1.log
