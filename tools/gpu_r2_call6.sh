#!/bin/bash
# Round 2, GPU call 6: ncu --set full captures of the decode kernels (stall reasons, pipes, memory), 1 vs 2 attention splits.
mkdir -p gpurun_out
T=r2f
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_kernel -s 10 -c 2 -o gpurun_out/prof_attn_decode_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled > gpurun_out/ncu_attn_$T.log 2>&1; tail -2 gpurun_out/ncu_attn_$T.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_kernel -s 10 -c 2 -o gpurun_out/prof_attn_decode_split2_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled --attn-splits 2 > gpurun_out/ncu_attn2_$T.log 2>&1; tail -2 gpurun_out/ncu_attn2_$T.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_tcgen05_kernel -s 30 -c 6 -o gpurun_out/prof_decode_gemm_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled > gpurun_out/ncu_gemm_$T.log 2>&1; tail -2 gpurun_out/ncu_gemm_$T.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"decode_residual_rmsnorm|argmax_advance|decode_embed" -s 4 -c 3 -o gpurun_out/prof_decode_small_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled > gpurun_out/ncu_small_$T.log 2>&1; tail -2 gpurun_out/ncu_small_$T.log | cut -c1-200
timeout 300 python tools/decode_ablate.py --mode tiled --quick > gpurun_out/ablate_${T}_tiled.json 2>&1; head -c 600 gpurun_out/ablate_${T}_tiled.json
ls -la gpurun_out/*.ncu-rep
