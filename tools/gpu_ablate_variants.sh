#!/bin/bash
mkdir -p gpurun_out
run() { timeout 150 python tools/decode_ablate.py > gpurun_out/decode_ablate_$1.json 2>&1; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/decode_ablate_$1.json"))
    print("$1", {k:(v if not isinstance(v,dict) else v["per_layer_us"]) for k,v in d.items()})
except Exception as e:
    print("$1 failed", e, open("gpurun_out/decode_ablate_$1.json").read()[-500:])
PY
}
run default
for v in "$@"; do DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_$v.so run $v; done
