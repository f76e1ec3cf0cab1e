#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attn_decode or decode" --timeout 60 2>&1 | tail -5
timeout 300 python tools/bench_ops.py decode 2>&1 | grep attn_decode | cut -c1-160
timeout 120 python tools/decode_ablate.py > gpurun_out/decode_ablate_$1.json 2>&1; python - <<PY
import json
d=json.load(open("gpurun_out/decode_ablate_$1.json"))
print({k:(v if not isinstance(v,dict) else v["per_layer_us"]) for k,v in d.items()})
PY
