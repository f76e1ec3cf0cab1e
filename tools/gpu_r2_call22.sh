#!/bin/bash
# 8 GPUs of one box: the headline workload (64 pages per rank) and BASELINE configs[4]'s shape (4 pages of 1960^2 per rank, 2048 tokens)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "headline rc=$?"; tail -1 gpurun_out/bench_r2_${N}gpu.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 2 --warmup 3 --page 1960 --batch 4 --new-tokens 2048 > gpurun_out/bench_r2_hires_${N}gpu.json 2> gpurun_out/bench_r2_hires_${N}gpu.err
echo "hires rc=$?"; tail -1 gpurun_out/bench_r2_hires_${N}gpu.json | cut -c1-300
python - <<PY
import json
for f in ("gpurun_out/bench_r2_${N}gpu.json", "gpurun_out/bench_r2_hires_${N}gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","n_gpus","ms_per_step","e2e","e2e_u8","clocks")})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_r2_${N}gpu.err | cut -c1-300
