#!/bin/bash
# Round 2, GPU call 1: hardware fundamentals, decode ablation sweep, suite, baseline bench, launch list at the headline
# config, vLLM comparison.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=r2a
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu -lcuda && timeout 300 /tmp/microbench > gpurun_out/microbench_$T.jsonl 2>&1
tail -8 gpurun_out/microbench_$T.jsonl
abl() { name=$1; shift; timeout 200 python tools/decode_ablate.py "$@" > gpurun_out/ablate_${T}_$name.json 2> gpurun_out/ablate_${T}_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ablate_${T}_$name.json"))
    print("$name", {k:(v if not isinstance(v,dict) else v["per_layer_us"]) for k,v in d.items()})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ablate_${T}_$name.err").read()[-400:])
PY
}
abl default
abl split2 --quick --attn-splits 2
abl split3 --quick --attn-splits 3
abl split4 --quick --attn-splits 4
DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_st2.so abl st2_split3 --quick --attn-splits 3
DOTS_B200_LIB=$PWD/dots_ocr_b200/build/variants/lib_st6.so abl st6_split1 --quick
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$T.log 2>&1; tail -1 gpurun_out/ncu_list_$T.log | cut -c1-200; wc -l gpurun_out/launches_$T.csv
export HF_HUB_OFFLINE=1 TRANSFORMERS_OFFLINE=1 VLLM_NO_USAGE_STATS=1 VLLM_DO_NOT_TRACK=1 TOKENIZERS_PARALLELISM=false
timeout 300 python tools/make_checkpoint_dir.py --preset full --flavour peaked --out /tmp/dots_full 2>&1 | tail -1
timeout 780 python tools/vllm_compare.py --dir /tmp/dots_full --impl vllm --pages 64 --new-tokens 512 > gpurun_out/vllm_$T.json 2> gpurun_out/vllm_$T.err
echo "vllm rc=$?"; tail -5 gpurun_out/vllm_$T.err | cut -c1-400
timeout 400 python tools/vllm_compare.py --dir /tmp/dots_full --impl ours --pages 64 --new-tokens 512 > gpurun_out/ours_$T.json 2> gpurun_out/ours_$T.err
echo "ours rc=$?"; tail -3 gpurun_out/ours_$T.err | cut -c1-300
python tools/vllm_compare.py --diff gpurun_out/vllm_$T.json gpurun_out/ours_$T.json | tee gpurun_out/vllm_diff_$T.json
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
