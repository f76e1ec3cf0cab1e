#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --batch 16 --steps 1 --warmup 1 --new-tokens 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_b16n8.json 2>/dev/null
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b16n8.csv \
    python bench.py --batch 16 --steps 1 --warmup 0 --new-tokens 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_b16n8.log 2>&1; wc -l gpurun_out/launches_b16n8.csv; tail -1 gpurun_out/ncu_list_b16n8.log | cut -c1-200
