#!/bin/bash
# Round 2, GPU call 4: kernel timelines of the decode step, ring-depth variants, per-kernel ncu list of one step.
mkdir -p gpurun_out
T=r2d
V=$PWD/dots_ocr_b200/build/variants
for m in tiled perop fused; do timeout 200 python tools/decode_timeline.py --mode $m > gpurun_out/timeline_${T}_$m.txt 2>&1; tail -42 gpurun_out/timeline_${T}_$m.txt | cut -c1-200; done
timeout 200 python tools/decode_timeline.py --mode tiled --graph > gpurun_out/timeline_${T}_tiled_graph.txt 2>&1; tail -12 gpurun_out/timeline_${T}_tiled_graph.txt | cut -c1-200
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
DOTS_B200_LIB=$V/lib_sw6.so abl tiled_sw6 --mode tiled
DOTS_B200_LIB=$V/lib_sw8.so abl tiled_sw8 --mode tiled
DOTS_B200_LIB=$V/lib_st4.so abl tiled_st4 --mode tiled --quick
DOTS_B200_LIB=$V/lib_st6.so abl tiled_st6 --mode tiled --quick
DOTS_B200_LIB=$V/lib_st2.so abl tiled_st2_split2 --mode tiled --attn-splits 2 --quick
DOTS_B200_LIB=$V/lib_dg2.so abl fused_dg2 --mode fused --quick
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/decode_step_${T}_tiled.csv \
    python tools/decode_step_profile.py --steps 3 --mode tiled > gpurun_out/decode_step_${T}_tiled.log 2>&1
python tools/decode_step_profile.py --summarise gpurun_out/decode_step_${T}_tiled.csv --steps 3 > gpurun_out/decode_traffic_${T}_tiled.json 2>&1; head -c 3000 gpurun_out/decode_traffic_${T}_tiled.json
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "rope or resize or processor" --timeout 200 2>&1 | tail -3
