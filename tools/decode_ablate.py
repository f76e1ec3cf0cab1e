#!/usr/bin/env python
"""In-situ cost of each decode-step component: time the captured decode graph with one op class stubbed out.
(Results of the stubbed runs are garbage; only the time difference against the full step matters.)

    python tools/decode_ablate.py [--batch 64] [--ctx 1881] [--steps 200]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import config, weights, ops  # noqa: E402
from dots_ocr_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1881)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--attn-splits", dest="attn_splits", type=int, default=0, help="override the flash-decoding split count of the plan")
    ap.add_argument("--quick", action="store_true", help="only the full step and the attention ablation")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = config.full()
    ck = weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev)
    eng = Engine(cfg, ck, dev)
    del ck
    B = args.batch
    ctx_max = (args.ctx + args.steps + 2 + 63) // 64 * 64
    kc, vc = eng._alloc_cache(B, ctx_max)
    kc.normal_(); vc.normal_()
    lens = torch.full((B,), args.ctx, device=dev, dtype=torch.int64)

    real = {n: getattr(ops, n) for n in ("gemm_skinny", "gemm_skinny_swiglu", "attn_decode_fused", "decode_residual_rmsnorm",
                                         "decode_embed_rmsnorm", "argmax_advance")}

    if args.attn_splits:
        plan0 = eng._decode_plan
        eng._decode_plan = lambda b: dict(plan0(b), attn=args.attn_splits)

    def run(stub=(), which_skinny=None):
        st = eng._new_decode_state(B, lens, kc, vc, ctx_max, args.steps + 2)
        st["last"].random_(0, 150000)
        for n, f in real.items():
            setattr(ops, n, f)
        for n in stub:
            if n == "gemm_skinny" and which_skinny is not None:
                orig = real["gemm_skinny"]

                def filt(x, w, splits=1, partial=None, out_bf16=None, bias=None, _o=orig, _w=which_skinny):
                    if w.shape == _w:
                        return None
                    return _o(x, w, splits, partial=partial, out_bf16=out_bf16, bias=bias)
                ops.gemm_skinny = filt
            else:
                setattr(ops, n, lambda *a, **k: None)
        eng._decode_step(st)
        torch.cuda.synchronize()
        g = ops.Graph()
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            with g:
                eng._decode_step(st)
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

    t = cfg.text
    H, I = t.hidden_size, t.intermediate_size
    qkv_n = (t.num_attention_heads + 2 * t.num_key_value_heads) * t.head_dim
    full = run()
    res = {"full_ms": round(full, 4)}
    cases = [("attention(+qkv finalize)", ("attn_decode_fused",), None)] if args.quick else None
    for name, stub, shape in cases or [("attention(+qkv finalize)", ("attn_decode_fused",), None),
                              ("qkv gemm", ("gemm_skinny",), (qkv_n, H)), ("o gemm", ("gemm_skinny",), (H, H)),
                              ("down gemm", ("gemm_skinny",), (H, I)), ("lm_head gemm", ("gemm_skinny",), (t.vocab_size, H)),
                              ("gate|up gemm + swiglu", ("gemm_skinny_swiglu",), None),
                              ("residual+rmsnorm finalize", ("decode_residual_rmsnorm",), None),
                              ("argmax", ("argmax_advance",), None),
                              ("all gemms", ("gemm_skinny", "gemm_skinny_swiglu"), None)]:
        ms = run(stub, shape)
        res[name] = {"without_ms": round(ms, 4), "cost_ms": round(full - ms, 4), "per_layer_us": round((full - ms) * 1e3 / t.num_hidden_layers, 2)}
    res["full_again_ms"] = round(run(), 4)
    w_bytes = eng.decode_weight_bytes()
    kv_bytes = 2 * t.num_hidden_layers * t.num_key_value_heads * t.head_dim * 2 * args.ctx * B
    res["roofline_ms_at_6485GBs"] = round((w_bytes + kv_bytes) / 6485.2e9 * 1e3, 4)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
