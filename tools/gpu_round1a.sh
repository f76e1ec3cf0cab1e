#!/bin/bash
# First GPU pass of the session: tc attention parity + micro-bench, headline bench, ncu launch list + full captures.
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
bash tests/run_gpu.sh tests/test_attn_tc_gpu.py
timeout 300 python tools/bench_ops.py attn > gpurun_out/bench_attn.log 2>&1; tail -8 gpurun_out/bench_attn.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
# launch list (cold-cache, serialised): one un-warmed step, 8 new tokens so the list stays short
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 0 --new-tokens 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1; tail -3 gpurun_out/ncu_list.log
# full captures of the dominant kernels (small batch: same kernels, same per-CTA behaviour)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 40 -c 4 -o gpurun_out/prof_gemm_r1 -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 4 -c 2 -o gpurun_out/prof_attn_r1 -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out
