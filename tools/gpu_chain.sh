#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "decode_chain" --timeout 60 2>&1 | tail -4 | cut -c1-300
timeout 150 python -m pytest tests/test_engine_gpu.py -x -q -m gpu --timeout 100 2>&1 | tail -4 | cut -c1-300
DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_chtime.so timeout 100 python tools/chain_timing.py 2>&1 | tail -3 | cut -c1-520
run() { timeout 150 python tools/decode_ablate.py > gpurun_out/decode_ablate_$1.json 2>&1; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/decode_ablate_$1.json"))
    print("$1", {k:(v if not isinstance(v,dict) else v["per_layer_us"]) for k,v in d.items()})
except Exception as e:
    print("$1 failed", open("gpurun_out/decode_ablate_$1.json").read()[-800:])
PY
}
run chain4
DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_ch8.so run chain8
