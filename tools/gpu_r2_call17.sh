#!/bin/bash
# max shared-memory carve-out on every decode kernel (co-residency of PDL neighbours): A/B against the previous build, with and without the
# 128-register attention; ring depths that let gate|up and down share an SM; the pipeline tests again with the whole log kept
mkdir -p gpurun_out
abl() { name=$1; lib=$2; shift; shift; DOTS_B200_LIB=$lib timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_r2o_$name.json 2> gpurun_out/ablate_r2o_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_r2o_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v.get("mode"))) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_r2o_$name.err").read()[-600:])
PY
}
V=dots_ocr_b200/build/variants
abl carve "" --mode tiled --quick
abl nocarve $V/lib_nocarve.so --mode tiled --quick
abl carve_nreg128 $V/lib_nreg128c.so --mode tiled --quick
abl carve_s444 "" --mode tiled --quick --stages 4,4,4
abl carve_s544 "" --mode tiled --quick --stages 5,4,4
abl carve_nreg128_s444 $V/lib_nreg128c.so --mode tiled --quick --stages 4,4,4
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_r2o_carve.txt 2>&1; head -12 gpurun_out/timeline_r2o_carve.txt | cut -c1-200
DOTS_B200_LIB=$V/lib_nreg128c.so timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_r2o_nreg128c.txt 2>&1; head -6 gpurun_out/timeline_r2o_nreg128c.txt | cut -c1-200
for t in tests/test_partition_gpu.py "tests/test_pipeline_gpu.py::test_pipeline_matches_generate" "tests/test_pipeline_gpu.py::test_pipeline_stop_ids_and_u8_pages"; do
  n=$(echo $t | tr '/:.' '___')
  timeout 300 python -X faulthandler -m pytest "$t" -x -v --timeout 250 > gpurun_out/pytest_$n.log 2>&1; echo "== $t rc=$?"; grep -n "Fatal\|Error\|error\|passed\|failed\|PASSED\|FAILED" gpurun_out/pytest_$n.log | head -12 | cut -c1-300
  grep -n "Current thread" -A12 gpurun_out/pytest_$n.log | head -30 | cut -c1-200
done
