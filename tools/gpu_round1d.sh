#!/bin/bash
mkdir -p gpurun_out
bash tests/run_gpu.sh tests/test_ops_gpu.py tests/test_attn_tc_gpu.py tests/test_engine_gpu.py || exit 1
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -5 gpurun_out/bench_d.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_d.json"))
print("bench_d", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline_decode"]["ms_per_decode_step"], d["roofline"]["achieved"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()})
PY
