#!/bin/bash
mkdir -p gpurun_out
set -x
bash tests/run_gpu.sh tests/test_ops_gpu.py tests/test_attn_tc_gpu.py tests/test_engine_gpu.py || exit 1
timeout 300 python tools/bench_ops.py skinny > gpurun_out/bench_ops_c.log 2>&1; cat gpurun_out/bench_ops_c.log | cut -c1-200
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; tail -c 1500 gpurun_out/bench_pdl.json; tail -5 gpurun_out/bench_pdl.err
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pdl --no-e2e > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; tail -c 1500 gpurun_out/bench_nopdl.json; tail -5 gpurun_out/bench_nopdl.err
