#!/bin/bash
# Round-end evidence: full GPU test suite, smoke, headline bench (with CPU baseline), reduced-batch bench + ncu launch list of the
# same command, full captures of the dominant kernels.
mkdir -p gpurun_out
T=$1
timeout 400 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -2 gpurun_out/bench_$T.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$T.json"))
print("bench", d["value"], "e2e", d["e2e"]["value"], d["ms_per_step"], "dec", d["roofline_decode"], "gemm", d["roofline"]["achieved"], d["roofline"]["frac"], {k:(v["ms"],v.get("tflops")) for k,v in d["kernels"].items()}, d["clocks"], d["cpu_baseline"]["value"], d["gpu_launches"])
PY
timeout 200 python tools/bench_ops.py gemm attn decode > gpurun_out/bench_ops_$T.jsonl 2>&1; tail -3 gpurun_out/bench_ops_$T.jsonl | cut -c1-200
# reduced-batch command, once plain (shares from CUDA events) and once under ncu (launch list)
timeout 300 python bench.py --batch 16 --steps 1 --warmup 1 --new-tokens 64 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16_$T.json 2>/dev/null
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --batch 16 --steps 1 --warmup 0 --new-tokens 64 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$T.log 2>&1; wc -l gpurun_out/launches_$T.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16|attn_fwd_tcgen05" -s 30 -c 6 -o gpurun_out/prof_prefill_$T -f \
    python bench.py --batch 8 --steps 1 --warmup 0 --new-tokens 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_prefill_$T.log 2>&1; tail -1 gpurun_out/ncu_prefill_$T.log | cut -c1-150
