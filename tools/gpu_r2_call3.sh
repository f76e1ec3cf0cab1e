#!/bin/bash
# Round 2, GPU call 3: decode layer variants over bulk-copied tiled operands (tests, ablation sweep, per-kernel profile), GPU resize,
# bench.
mkdir -p gpurun_out
T=r2c
timeout 900 python -m pytest tests/test_decode_fused_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_zz_stop_ids_gpu.py tests/test_zzz_continuous_gpu.py -x -q --timeout 300 2>&1 | tail -15
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
V=$PWD/dots_ocr_b200/build/variants
abl tiled --mode tiled
abl fused --mode fused
abl perop --mode perop --quick
abl tiled_split2 --mode tiled --attn-splits 2 --quick
abl fused_split2 --mode fused --attn-splits 2 --quick
DOTS_B200_LIB=$V/lib_sw6.so abl tiled_sw6 --mode tiled --quick
DOTS_B200_LIB=$V/lib_sw8.so abl tiled_sw8 --mode tiled --quick
DOTS_B200_LIB=$V/lib_st4.so abl tiled_st4 --mode tiled --quick
DOTS_B200_LIB=$V/lib_st2.so abl tiled_st2_split2 --mode tiled --attn-splits 2 --quick
DOTS_B200_LIB=$V/lib_dg2.so abl fused_dg2 --mode fused --quick
DOTS_B200_LIB=$V/lib_dg3.so abl fused_dg3 --mode fused --quick
for m in tiled fused; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dots --csv --log-file gpurun_out/decode_step_${T}_$m.csv \
    python tools/decode_step_profile.py --steps 3 --mode $m > gpurun_out/decode_step_${T}_$m.log 2>&1
python tools/decode_step_profile.py --summarise gpurun_out/decode_step_${T}_$m.csv --steps 3 > gpurun_out/decode_traffic_${T}_$m.json 2>&1; head -c 2500 gpurun_out/decode_traffic_${T}_$m.json; rm -f gpurun_out/decode_step_${T}_$m.csv.bak
done
timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-300; tail -3 gpurun_out/bench_$T.err
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --decode-mode fused --no-e2e > gpurun_out/bench_${T}_fused.json 2> gpurun_out/bench_${T}_fused.err; tail -1 gpurun_out/bench_${T}_fused.json | cut -c1-200
