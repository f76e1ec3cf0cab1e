#!/bin/bash
# Round 2, GPU call 10: per-launch ring depths of the decode GEMMs.
mkdir -p gpurun_out
T=r2j
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -x -q --timeout 300 2>&1 | tail -3
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
abl s454 --mode tiled
abl s444 --mode tiled --stages 4,4,4 --quick
abl s464 --mode tiled --stages 4,6,4 --quick
abl s354 --mode tiled --stages 3,5,4 --quick
abl s654 --mode tiled --stages 6,5,4 --quick
abl s456 --mode tiled --stages 4,5,6 --quick
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_${T}_tiled.txt 2>&1; head -12 gpurun_out/timeline_${T}_tiled.txt | cut -c1-200
