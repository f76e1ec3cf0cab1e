#!/bin/bash
mkdir -p gpurun_out
python - <<PY
import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("no cpu.max", e)
PY
for f in tests/test_ops_gpu.py tests/test_attn_tc_gpu.py tests/test_engine_gpu.py; do timeout 200 python -m pytest $f -x -q -m gpu --timeout 120 2>&1 | tail -2; done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full_$1.json 2> gpurun_out/bench_full_$1.err; tail -3 gpurun_out/bench_full_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full_$1.json"))
print("bench", d["value"], "e2e", d["e2e"]["value"], d["ms_per_step"], "dec", d["roofline_decode"]["ms_per_decode_step"], "gemm", d["roofline"]["achieved"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()}, d["clocks"], d["cpu_baseline"])
PY
