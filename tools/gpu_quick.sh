#!/bin/bash
# usage: tools/gpu_quick.sh <tag> [bench args...]   -- tc-attention tests + attention micro-bench + short bench
mkdir -p gpurun_out
tag=$1; shift
bash tests/run_gpu.sh tests/test_attn_tc_gpu.py tests/test_engine_gpu.py || exit 1
timeout 300 python tools/bench_ops.py attn > gpurun_out/bench_attn_$tag.log 2>&1; grep '"tc"' gpurun_out/bench_attn_$tag.log | cut -c1-200
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", d["value"], d["ms_per_step"], d["roofline_decode"]["ms_per_decode_step"], d["roofline"]["achieved"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()})
PY
