#!/bin/bash
# 2 GPUs: bench under torchrun with NCCL_DEBUG=INFO (the JSON must be the last stdout line), per-step times of the host-buffer legs
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_2gpu_b.out 2> gpurun_out/bench_r2_2gpu_b.err
echo "rc=$?"; echo "--- last stdout line:"; tail -1 gpurun_out/bench_r2_2gpu_b.out | cut -c1-600
python - <<'PY'
import json
l=open("gpurun_out/bench_r2_2gpu_b.out").read().strip().splitlines()
print("stdout lines:", len(l))
d=json.loads(l[-1]); print({k:d.get(k) for k in ("value","ms_per_step","e2e","e2e_u8","phases_ms","clocks")})
PY
ls gpurun_out/ | grep -i nccl | head; grep -c NCCL gpurun_out/bench_r2_2gpu_b.err
