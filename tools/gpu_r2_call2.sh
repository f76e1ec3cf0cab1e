#!/bin/bash
# Round 2, GPU call 2: 2-D tensor-map streaming rates, the new decode kernels (tests + ablation), parity on the bench config,
# quick bench in the new JSON format, vLLM comparison retry.
mkdir -p gpurun_out
T=r2b
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu -lcuda && timeout 200 /tmp/microbench quick > gpurun_out/microbench_$T.jsonl 2>&1
grep stream2d gpurun_out/microbench_$T.jsonl | head -40
timeout 600 python -m pytest tests/test_decode_fused_gpu.py tests/test_zz_stop_ids_gpu.py tests/test_engine_gpu.py -x -q --timeout 300 2>&1 | tail -15
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "argmax or decode" --timeout 300 2>&1 | tail -3
abl() { name=$1; shift; timeout 250 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v.get("per_layer_us", v)) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-600:])
PY
}
abl fused --fused 1
abl fused_split1 --fused 1 --attn-splits 1 --quick
abl fused_split4 --fused 1 --attn-splits 4 --quick
abl perop --fused 0 --quick
timeout 900 python -m pytest tests/test_bench_config_gpu.py -x -q -s --timeout 800 2>&1 | grep -v "^$" | tail -40
timeout 400 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-1500; tail -3 gpurun_out/bench_$T.err
export HF_HUB_OFFLINE=1 TRANSFORMERS_OFFLINE=1 VLLM_NO_USAGE_STATS=1 VLLM_DO_NOT_TRACK=1 TOKENIZERS_PARALLELISM=false
timeout 300 python tools/make_checkpoint_dir.py --preset full --flavour peaked --out /tmp/dots_full 2>&1 | tail -1
timeout 720 python tools/vllm_compare.py --dir /tmp/dots_full --impl vllm --pages 64 --new-tokens 512 > gpurun_out/vllm_$T.json 2> gpurun_out/vllm_$T.err
echo "vllm rc=$?"; grep -v "^$" gpurun_out/vllm_$T.err | tail -12 | cut -c1-400
timeout 300 python tools/vllm_compare.py --dir /tmp/dots_full --impl ours --pages 64 --new-tokens 512 > gpurun_out/ours_$T.json 2> gpurun_out/ours_$T.err
echo "ours rc=$?"
python tools/vllm_compare.py --diff gpurun_out/vllm_$T.json gpurun_out/ours_$T.json 2>&1 | tail -2 | tee gpurun_out/vllm_diff_$T.json
python - <<PY
import json
for n in ("vllm", "ours"):
    p = f"gpurun_out/{n}_$T.json"
    try:
        d = json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1]); d.pop("ids", None)
        json.dump(d, open(p, "w")); print(n, d)
    except Exception as e:
        print(n, "no result:", e)
PY
