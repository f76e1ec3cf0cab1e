#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/bench_ops.py gemm 2>&1 | grep '"gemm"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['M'], d['N'], d['K'], d['epi'], 'single', d['tflops'], 'pair', d['pair_tflops'], 'cublas', d['cublas_tflops'])"
for gp in 0 1; do
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --gemm-pair $gp > gpurun_out/bench_pair$gp.json 2> gpurun_out/bench_pair$gp.err; tail -2 gpurun_out/bench_pair$gp.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_pair$gp.json"))
print("pair$gp", d["value"], d["ms_per_step"], d["roofline_decode"]["ms_per_decode_step"], d["roofline"]["achieved"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()}, d["clocks"])
PY
done
