#!/bin/bash
# Round 2, GPU call 9: two consumer groups in the decode attention: tests, ablation, timeline.
mkdir -p gpurun_out
T=r2i
V=$PWD/dots_ocr_b200/build/variants
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_zz_stop_ids_gpu.py tests/test_zzz_continuous_gpu.py -x -q --timeout 300 2>&1 | tail -4
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
DOTS_B200_LIB=$V/lib_g1.so abl tiled_g1 --mode tiled --quick
DOTS_B200_LIB=$V/lib_g2st6.so abl tiled_g2st6 --mode tiled --quick
DOTS_B200_LIB=$V/lib_g2st2.so abl tiled_g2st2 --mode tiled --quick
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_${T}_tiled.txt 2>&1; head -12 gpurun_out/timeline_${T}_tiled.txt | cut -c1-200
timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-300; tail -3 gpurun_out/bench_$T.err
