#!/bin/bash
# page pipeline on SM partitions: correctness (tiny config) + full-size sequential vs pipelined; decode attention capped at 128 registers
# (co-resident with the q|k|v GEMM's CTAs -> earlier KV prefetch): ablation + timeline against the product build
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_pipeline_gpu.py tests/test_partition_gpu.py -x -q --timeout 300 2>&1 | tail -15
timeout 400 python tools/pipeline_bench.py --first 96 --batches 3 > gpurun_out/pipeline_bench_96.jsonl 2> gpurun_out/pipeline_bench_96.err; echo "pipeline_bench rc=$?"; tail -2 gpurun_out/pipeline_bench_96.jsonl | cut -c1-900; tail -4 gpurun_out/pipeline_bench_96.err | cut -c1-600
abl() { name=$1; lib=$2; shift; shift; DOTS_B200_LIB=$lib timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_r2n_$name.json 2> gpurun_out/ablate_r2n_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_r2n_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_r2n_$name.err").read()[-600:])
PY
}
V=dots_ocr_b200/build/variants
abl base "" --mode tiled --quick
abl nreg128 $V/lib_nreg128.so --mode tiled --quick
abl base2 "" --mode tiled --quick
abl nreg128b $V/lib_nreg128.so --mode tiled --quick
DOTS_B200_LIB=$V/lib_nreg128.so timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_r2n_nreg128.txt 2>&1; head -12 gpurun_out/timeline_r2n_nreg128.txt | cut -c1-200
DOTS_B200_LIB=$V/lib_nreg128.so timeout 300 python -m pytest tests/test_decode_fused_gpu.py -x -q --timeout 250 2>&1 | tail -3
