#!/bin/bash
# Round 2, GPU call 5: merged finalize (rendezvous inside o_proj / down_proj): tests, timeline, ablation, bench.
mkdir -p gpurun_out
T=r2e
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_engine_gpu.py -x -q --timeout 300 2>&1 | tail -6
for m in tiled; do timeout 200 python tools/decode_timeline.py --mode $m > gpurun_out/timeline_${T}_$m.txt 2>&1; head -16 gpurun_out/timeline_${T}_$m.txt | cut -c1-220; done
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
abl tiled --mode tiled
abl tiled7 --mode tiled7 --quick
timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-400; tail -3 gpurun_out/bench_$T.err
