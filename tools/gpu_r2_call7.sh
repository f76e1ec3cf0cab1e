#!/bin/bash
# Round 2, GPU call 7: lean decode-attention consumer loop: tests, ablation (ring depths, splits), timeline.
mkdir -p gpurun_out
T=r2g
V=$PWD/dots_ocr_b200/build/variants
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -x -q --timeout 300 2>&1 | tail -4
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
abl tiled --mode tiled
abl tiled_split2 --mode tiled --attn-splits 2 --quick
DOTS_B200_LIB=$V/lib_st4.so abl tiled_st4 --mode tiled --quick
DOTS_B200_LIB=$V/lib_st6.so abl tiled_st6 --mode tiled --quick
DOTS_B200_LIB=$V/lib_sw6.so abl tiled_sw6 --mode tiled --quick
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_${T}_tiled.txt 2>&1; head -12 gpurun_out/timeline_${T}_tiled.txt | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_decode_kernel -s 10 -c 1 -o gpurun_out/prof_attn_decode_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled > gpurun_out/ncu_attn_$T.log 2>&1; tail -1 gpurun_out/ncu_attn_$T.log | cut -c1-200
