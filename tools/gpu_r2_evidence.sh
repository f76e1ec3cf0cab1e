#!/bin/bash
# Round 2 evidence pass: full GPU suite, final bench, decode ablation / timeline / per-kernel ncu list, headline launch list, ncu --set full
# captures (decode kernels, HBM-bound elementwise kernels, prefill kernels), op micro-benchmarks, vLLM comparison.
# Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
T=r2final
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 2>&1 | tail -4 | tee gpurun_out/pytest_$T.txt
grep -q "passed" gpurun_out/pytest_$T.txt || { echo "GPU suite not green: stopping"; exit 1; }
grep -q "failed\|error" gpurun_out/pytest_$T.txt && { echo "GPU suite not green: stopping"; exit 1; }
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -1 gpurun_out/bench_$T.json | cut -c1-500; tail -2 gpurun_out/bench_$T.err
timeout 250 python tools/decode_ablate.py --mode tiled > gpurun_out/ablate_$T.json 2> gpurun_out/ablate_$T.err; python -c "
import json; d=json.load(open('gpurun_out/ablate_$T.json')); print({k:(v if not isinstance(v,dict) else v.get('per_layer_us', v.get('mode'))) for k,v in d.items()})"
timeout 200 python tools/decode_timeline.py --mode tiled > gpurun_out/timeline_$T.txt 2>&1; head -11 gpurun_out/timeline_$T.txt | cut -c1-180
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/decode_step_$T.csv \
    python tools/decode_step_profile.py --steps 3 --mode tiled > gpurun_out/decode_step_$T.log 2>&1
python tools/decode_step_profile.py --summarise gpurun_out/decode_step_$T.csv --steps 3 > gpurun_out/decode_traffic_$T.json 2>&1; tail -4 gpurun_out/decode_traffic_$T.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"attn_decode_kernel|gemm_bf16_tcgen05_kernel|decode_residual_rmsnorm|argmax_advance|decode_embed" -s 8 -c 9 -o gpurun_out/prof_decode_$T -f \
    python tools/decode_step_profile.py --steps 1 --mode tiled > gpurun_out/ncu_decode_$T.log 2>&1; tail -1 gpurun_out/ncu_decode_$T.log | cut -c1-160
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"rmsnorm_kernel|llm_rope_append|cast_pad|patchify|resample|vit_rope_table|embed_scatter|layernorm" -c 12 -o gpurun_out/prof_elementwise_$T -f \
    python bench.py --batch 8 --new-tokens 4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_elem_$T.log 2>&1; tail -1 gpurun_out/ncu_elem_$T.log | cut -c1-160
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"gemm2_bf16|attn_fwd_tcgen05" -s 6 -c 6 -o gpurun_out/prof_prefill_$T -f \
    python bench.py --batch 8 --new-tokens 4 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_prefill_$T.log 2>&1; tail -1 gpurun_out/ncu_prefill_$T.log | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$T.log 2>&1; wc -l gpurun_out/launches_$T.csv
timeout 400 python tools/bench_ops.py attn decode_gemm > gpurun_out/bench_ops_$T.jsonl 2> gpurun_out/bench_ops_$T.err; cat gpurun_out/bench_ops_$T.jsonl | cut -c1-220
export HF_HUB_OFFLINE=1 TRANSFORMERS_OFFLINE=1 VLLM_NO_USAGE_STATS=1 VLLM_DO_NOT_TRACK=1 TOKENIZERS_PARALLELISM=false
timeout 300 python tools/make_checkpoint_dir.py --preset full --flavour peaked --out /tmp/dots_full 2>&1 | tail -1
timeout 600 python tools/vllm_compare.py --dir /tmp/dots_full --impl vllm --pages 64 --new-tokens 512 > gpurun_out/vllm_$T.json 2> gpurun_out/vllm_$T.err; echo "vllm rc=$?"
timeout 300 python tools/vllm_compare.py --dir /tmp/dots_full --impl ours --pages 64 --new-tokens 512 > gpurun_out/ours_$T.json 2> gpurun_out/ours_$T.err; echo "ours rc=$?"; tail -2 gpurun_out/ours_$T.err | cut -c1-300
python tools/vllm_compare.py --diff gpurun_out/vllm_$T.json gpurun_out/ours_$T.json 2>&1 | tail -1 | tee gpurun_out/vllm_diff_$T.json
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
ls -la gpurun_out/*$T*.ncu-rep
