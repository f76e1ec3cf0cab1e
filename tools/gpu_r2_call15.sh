#!/bin/bash
# SM-partition (green context) go / no-go: basic test, then the probe at three splits; GEMM-vs-cuBLAS rows for the record
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_partition_gpu.py -x -q -s --timeout 200 2>&1 | tail -15
for f in 96 104 88; do
  timeout 300 python tools/partition_probe.py --first $f > gpurun_out/partition_probe_$f.jsonl 2> gpurun_out/partition_probe_$f.err
  echo "== first=$f rc=$?"; tail -1 gpurun_out/partition_probe_$f.jsonl | cut -c1-1500; tail -3 gpurun_out/partition_probe_$f.err | cut -c1-400
done
timeout 300 python tools/bench_ops.py gemm > gpurun_out/bench_ops_r2_gemm.jsonl 2> gpurun_out/bench_ops_r2_gemm.err; tail -12 gpurun_out/bench_ops_r2_gemm.jsonl | cut -c1-300
