#!/bin/bash
# First GPU call of round 2: confirm the state round 1 ended in, then the vLLM comparison that could not run in round 1
# (SURVEY §8f N4).  Usage (build container):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r2a'
mkdir -p gpurun_out
T=${1:-r2a}
timeout 500 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
# the cases that had not run on hardware when round 1 ended: XPASS = the slot backend works, xfail = read the reason
timeout 300 python -m pytest tests/test_zzz_continuous_gpu.py tests/test_zz_vllm_golden_gpu.py -q -rxX --timeout 200 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-300
# vLLM 0.22 DotsOCRForCausalLM on the same synthetic parameters; each arm is its own process
timeout 300 python tools/make_checkpoint_dir.py --preset full --flavour peaked --out /tmp/dots_full 2>&1 | tail -1
timeout 900 python tools/vllm_compare.py --dir /tmp/dots_full --impl vllm --pages 64 --new-tokens 512 > gpurun_out/vllm_$T.json 2> gpurun_out/vllm_$T.err
tail -3 gpurun_out/vllm_$T.err | cut -c1-300
timeout 600 python tools/vllm_compare.py --dir /tmp/dots_full --impl ours --pages 64 --new-tokens 512 > gpurun_out/ours_$T.json 2> gpurun_out/ours_$T.err
tail -3 gpurun_out/ours_$T.err | cut -c1-300
python tools/vllm_compare.py --diff gpurun_out/vllm_$T.json gpurun_out/ours_$T.json | tee gpurun_out/vllm_diff_$T.json
# the id lists are large; keep the timings only
python - <<PY
import json
for n in ("vllm", "ours"):
    p = f"gpurun_out/{n}_$T.json"
    try:
        d = json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1]); d.pop("ids", None)
        json.dump(d, open(p, "w"))
    except Exception as e:
        print(n, "no result:", e)
PY
