#!/bin/bash
mkdir -p gpurun_out
set -x
bash tests/run_gpu.sh tests/test_ops_gpu.py tests/test_attn_tc_gpu.py tests/test_engine_gpu.py
timeout 300 python tools/bench_ops.py gemm skinny decode > gpurun_out/bench_ops_b.log 2>&1; cat gpurun_out/bench_ops_b.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full_b.json 2> gpurun_out/bench_full_b.err; tail -c 2500 gpurun_out/bench_full_b.json; tail -5 gpurun_out/bench_full_b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05 -s 4 -c 2 -o gpurun_out/prof_attn_tc_r1 -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_attn_tc.log 2>&1; tail -3 gpurun_out/ncu_attn_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_decode_kernel|gemm_bf16_tcgen05_kernel<64" -s 30 -c 8 -o gpurun_out/prof_decode_r1 -f \
    python bench.py --batch 64 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_decode.log 2>&1; tail -3 gpurun_out/ncu_decode.log
