#!/bin/bash
# Round-1 evidence pass: launch list (shares) + full captures of the dominant kernels.  Numbers printed by runs under ncu
# are never bench values.
mkdir -p gpurun_out
T=$1
# 1) launch list, application replay (one pass, no per-kernel memory save/restore)
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --replay-mode application -c 6000 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 0 --new-tokens 16 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$T.log 2>&1; tail -2 gpurun_out/ncu_list_$T.log | cut -c1-300; wc -l gpurun_out/launches_$T.csv
# 2) full captures, prefill-side kernels at batch 8 (same kernels / per-CTA behaviour, short run)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 40 -c 6 -o gpurun_out/prof_gemm_$T -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gemm_$T.log 2>&1; tail -1 gpurun_out/ncu_gemm_$T.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05 -s 4 -c 2 -o gpurun_out/prof_attn_$T -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_attn_$T.log 2>&1; tail -1 gpurun_out/ncu_attn_$T.log | cut -c1-200
# 3) decode-side kernels at batch 64
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_decode_kernel|gemm_bf16_tcgen05_kernel<64|decode_residual" -s 20 -c 10 -o gpurun_out/prof_decode_$T -f \
    python bench.py --batch 64 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_decode_$T.log 2>&1; tail -1 gpurun_out/ncu_decode_$T.log | cut -c1-200
ls -la gpurun_out | tail -12
