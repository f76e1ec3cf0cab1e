#!/bin/bash
mkdir -p gpurun_out
T=r2m
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_zz_stop_ids_gpu.py -x -q --timeout 300 2>&1 | tail -3
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
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_${T}_tiled.txt 2>&1; head -12 gpurun_out/timeline_${T}_tiled.txt | cut -c1-200
