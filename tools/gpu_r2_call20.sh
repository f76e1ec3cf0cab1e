#!/bin/bash
# dependency counters between the kernels of the decode step: correctness on the tiny config, then the step time with and without
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_decode_fused_gpu.py -x -q --timeout 300 2>&1 | tail -6
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_r2q_$name.json 2> gpurun_out/ablate_r2q_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_r2q_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_r2q_$name.err").read()[-900:])
PY
}
abl deps2 --mode tiled --quick --deps 1
abl nodeps2 --mode tiled --quick --deps 0
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_r2r_deps.txt 2>&1; head -12 gpurun_out/timeline_r2r_deps.txt | cut -c1-200
timeout 200 python tools/decode_timeline.py --mode tiled --graph > gpurun_out/timeline_r2r_deps_graph.txt 2>&1; head -12 gpurun_out/timeline_r2r_deps_graph.txt | cut -c1-200
python - <<'PY'
import torch
from dots_ocr_b200 import config, weights
from dots_ocr_b200.engine import Engine
import subprocess, sys
PY
DOTS_NO_DEPS=1 timeout 200 python - <<'PY' > gpurun_out/timeline_r2r_nodeps_graph.txt 2>&1
import sys, runpy
sys.argv = ["tools/decode_timeline.py", "--mode", "tiled", "--graph", "--deps", "0"]
runpy.run_path("tools/decode_timeline.py", run_name="__main__")
PY
head -10 gpurun_out/timeline_r2r_nodeps_graph.txt | cut -c1-200
