#!/bin/bash
# Round 2, GPU call 11: full GPU suite, smoke, bench, serving benchmark (mixed lengths), hi-res config.
mkdir -p gpurun_out
T=r2k
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-600; tail -3 gpurun_out/bench_$T.err
timeout 600 python tools/serve_bench.py --pages 192 --min-new 100 --max-new 2000 --threads 64 > gpurun_out/serve_$T.json 2> gpurun_out/serve_$T.err; tail -1 gpurun_out/serve_$T.json | cut -c1-900; tail -3 gpurun_out/serve_$T.err
timeout 500 python bench.py --page 1960 --batch 4 --new-tokens 2048 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${T}_hires.json 2> gpurun_out/bench_${T}_hires.err; tail -1 gpurun_out/bench_${T}_hires.json | cut -c1-400; tail -2 gpurun_out/bench_${T}_hires.err
