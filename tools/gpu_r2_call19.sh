#!/bin/bash
# resnorm early trigger A/B; then the whole GPU suite, smoke and the default bench on the product build
mkdir -p gpurun_out
abl() { name=$1; lib=$2; shift; shift; DOTS_B200_LIB=$lib timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_r2p_$name.json 2> gpurun_out/ablate_r2p_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_r2p_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_r2p_$name.err").read()[-600:])
PY
}
V=dots_ocr_b200/build/variants
abl base "" --mode tiled --quick
abl rnearly $V/lib_rnearly.so --mode tiled --quick
abl base2 "" --mode tiled --quick
abl rnearly2 $V/lib_rnearly.so --mode tiled --quick
DOTS_B200_LIB=$V/lib_rnearly.so timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_r2p_rnearly.txt 2>&1; head -10 gpurun_out/timeline_r2p_rnearly.txt | cut -c1-200
DOTS_B200_LIB=$V/lib_rnearly.so timeout 300 python -m pytest tests/test_decode_fused_gpu.py tests/test_zz_stop_ids_gpu.py -x -q --timeout 250 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","e2e","e2e_u8","ids_checksum","clocks","cpu_baseline")})
print(d["roofline"])
PY
