#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_ops_gpu.py -x -q -m gpu --timeout 100 2>&1 | tail -2
timeout 200 python tools/bench_ops.py gemm 2>&1 | grep '"gemm"' | cut -c1-200
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; tail -3 gpurun_out/bench_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json"))
print("$1", d["value"], d["ms_per_step"], d["roofline_decode"]["ms_per_decode_step"], d["roofline"]["achieved"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()})
PY
