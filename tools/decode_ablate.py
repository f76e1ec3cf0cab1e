#!/usr/bin/env python
"""In-situ cost of each decode-step component: time the captured decode graph with one op class stubbed out.
(Results of the stubbed runs are garbage; only the time difference against the full step matters.)

    python tools/decode_ablate.py [--batch 64] [--ctx 1881] [--steps 200] [--fused 0|1] [--attn-splits N] [--quick]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import config, weights, ops  # noqa: E402
from dots_ocr_b200.engine import Engine  # noqa: E402

OPS = ("decode_gemm_partial", "decode_gemm_swiglu", "decode_gemm_head", "gemm_skinny", "gemm_skinny_swiglu", "attn_decode_fused", "attn_decode_qkv", "decode_residual_rmsnorm", "decode_embed_rmsnorm",
       "argmax_advance", "decode_gemm_qkv", "decode_gemm_resnorm")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1881)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--mode", default="tiled", choices=["tiled", "fused", "perop"], help="decode layer variant (Engine.decode_mode)")
    ap.add_argument("--attn-splits", dest="attn_splits", type=int, default=0, help="override the flash-decoding split count of the plan")
    ap.add_argument("--no-cluster", action="store_true", help="combine kernel instead of the cluster merge for 2..4 attention splits")
    ap.add_argument("--quick", action="store_true", help="only the full step and the attention ablation")
    ap.add_argument("--stages", default=None, help="ring depths partial,swiglu,head of the decode GEMMs (e.g. 4,5,4)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = config.full()
    ck = weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev)
    eng = Engine(cfg, ck, dev)
    eng.decode_mode = args.mode
    del ck
    if args.no_cluster:
        ops.set_decode_cluster(False)
    if args.stages:
        ops.set_decode_stages(*[int(x) for x in args.stages.split(",")])
    B = args.batch
    ctx_max = (args.ctx + args.steps + 2 + 63) // 64 * 64
    kc, vc = eng._alloc_cache(B, ctx_max)
    kc.normal_(); vc.normal_()
    lens = torch.full((B,), args.ctx, device=dev, dtype=torch.int64)
    eng.attn_splits = args.attn_splits
    real = {n: getattr(ops, n) for n in OPS}
    t = cfg.text
    H, I = t.hidden_size, t.intermediate_size
    qkv_n = (t.num_attention_heads + 2 * t.num_key_value_heads) * t.head_dim

    def run(stub=()):
        """stub: iterable of (op name, weight shape or None)."""
        st = eng._new_decode_state(B, lens, kc, vc, ctx_max, args.steps + 2)
        st["last"].random_(0, 150000)
        for n, f in real.items():
            setattr(ops, n, f)
        for n, shape in stub:
            if shape is None:
                setattr(ops, n, lambda *a, **k: None)
            else:
                def filt(*a, _o=real[n], _s=tuple(shape), **k):
                    if tuple(a[1].shape) == _s:
                        return None
                    return _o(*a, **k)
                setattr(ops, n, filt)
        eng._decode_step(st)
        torch.cuda.synchronize()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            g = ops.capture(lambda: eng._decode_step(st))
            for _ in range(5):
                g.launch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cap)
            for _ in range(args.steps - 10):
                g.launch()
            e1.record(cap)
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        for n, f in real.items():
            setattr(ops, n, f)
        return e0.elapsed_time(e1) / (args.steps - 10)

    full = run()
    pl = eng._decode_plan(B)
    L0 = eng.t_layers[0]
    res = {"full_ms": round(full, 4), "plan": {k: v for k, v in pl.items()}}
    if pl["mode"] == "fused":
        cases = [("attention(+rope, append)", [("attn_decode_qkv", None)]),
                 ("qkv gemm (cluster)", [("decode_gemm_qkv", None)]),
                 ("o gemm + residual + norm (cluster)", [("decode_gemm_resnorm", L0["o_t"].shape)]),
                 ("down gemm + residual + norm (cluster)", [("decode_gemm_resnorm", L0["down_t"].shape)]),
                 ("gate|up gemm + swiglu", [("decode_gemm_swiglu", None)]),
                 ("lm_head gemm", [("decode_gemm_head", None)]),
                 ("argmax", [("argmax_advance", None)]),
                 ("all gemms", [("decode_gemm_qkv", None), ("decode_gemm_resnorm", None), ("decode_gemm_swiglu", None), ("decode_gemm_head", None)])]
    elif pl["mode"] == "tiled":
        cases = [("attention(+qkv finalize)", [("attn_decode_fused", None)]),
                 ("qkv gemm", [("decode_gemm_partial", L0["qkv_w_t"].shape)]), ("o gemm", [("decode_gemm_partial", L0["o_t"].shape)]),
                 ("down gemm", [("decode_gemm_partial", L0["down_t"].shape)]), ("lm_head gemm", [("decode_gemm_head", None)]),
                 ("gate|up gemm + swiglu", [("decode_gemm_swiglu", None)]),
                 ("residual+rmsnorm finalize", [("decode_residual_rmsnorm", None)]),
                 ("argmax", [("argmax_advance", None)]),
                 ("all gemms", [("decode_gemm_partial", None), ("decode_gemm_swiglu", None), ("decode_gemm_head", None)])]
    else:
        cases = [("attention(+qkv finalize)", [("attn_decode_fused", None)]),
                 ("qkv gemm", [("gemm_skinny", (qkv_n, H))]), ("o gemm", [("gemm_skinny", (H, H))]),
                 ("down gemm", [("gemm_skinny", (H, I))]), ("lm_head gemm", [("gemm_skinny", (t.vocab_size, H))]),
                 ("gate|up gemm + swiglu", [("gemm_skinny_swiglu", None)]),
                 ("residual+rmsnorm finalize", [("decode_residual_rmsnorm", None)]),
                 ("argmax", [("argmax_advance", None)]),
                 ("all gemms", [("gemm_skinny", None), ("gemm_skinny_swiglu", None)])]
    if args.quick:
        cases = cases[:1]
    for name, stub in cases:
        ms = run(stub)
        res[name] = {"without_ms": round(ms, 4), "cost_ms": round(full - ms, 4), "per_layer_us": round((full - ms) * 1e3 / t.num_hidden_layers, 2)}
    res["full_again_ms"] = round(run(), 4)
    w_bytes = eng.decode_weight_bytes()
    kv_bytes = 2 * t.num_hidden_layers * t.num_key_value_heads * t.head_dim * 2 * args.ctx * B
    res["roofline_ms_at_6485GBs"] = round((w_bytes + kv_bytes) / 6485.2e9 * 1e3, 4)
    res["frac_of_roofline"] = round(res["roofline_ms_at_6485GBs"] / full, 4)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
