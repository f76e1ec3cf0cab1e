#!/usr/bin/env python
"""Timeline of ONE decode step from the kernels' own timestamps (%globaltimer): for every kernel of the step, when its first CTA
started, when the dependency wait released it, when the first operand stage had landed, when the last one was consumed and when
the last CTA left.  Answers "where do the 70 us per layer go": kernel boundaries, prologues, streaming, epilogues.

    python tools/decode_timeline.py [--mode tiled|fused|perop] [--batch 64] [--ctx 1881] [--layers 3] [--graph]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import config, weights, ops  # noqa: E402
from dots_ocr_b200.engine import Engine  # noqa: E402

KID = {20: "attn_decode", 30: "residual_rmsnorm", 40: "cluster_gemm_qkv", 41: "cluster_gemm_resnorm", 105: "gemm_partial(F32_T)",
       106: "gemm_head(BF16_T)", 107: "gemm_swiglu(SWIGLU_T)"}
POINT = {0: "start", 1: "dep_released", 5: "prologue_done", 2: "first_stage", 3: "last_consumed", 4: "end", 6: "rendezvous_passed", 7: "finalize_done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="tiled")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1881)
    ap.add_argument("--attn-splits", dest="attn_splits", type=int, default=0)
    ap.add_argument("--graph", action="store_true", help="time a CUDA-graph replay of the step (default: eager launches with PDL)")
    ap.add_argument("--show-layers", dest="show", type=int, default=2, help="print the kernels of this many layers (from layer 3 on)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = config.full()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev), dev)
    eng.decode_mode = a.mode
    eng.attn_splits = a.attn_splits
    B = a.batch
    ctx_max = (a.ctx + 16 + 63) // 64 * 64
    kc, vc = eng._alloc_cache(B, ctx_max)
    kc.normal_(); vc.normal_()
    lens = torch.full((B,), a.ctx, device=dev, dtype=torch.int64)
    st = eng._new_decode_state(B, lens, kc, vc, ctx_max, 16)
    st["last"].random_(0, 150000)
    for _ in range(3):
        eng._decode_step(st)
    side = torch.cuda.Stream(device=dev)           # stream capture is not allowed on the legacy default stream
    torch.cuda.synchronize()
    cap = 400000
    buf = torch.zeros(2 + 3 * cap, device=dev, dtype=torch.int64)
    buf[1] = cap
    torch.cuda.synchronize()
    ops.debug_set_trace(buf)
    if a.graph:
        # the trace pointer is a kernel parameter: capture with it armed
        with torch.cuda.stream(side):
            g = ops.capture(lambda: eng._decode_step(st))
            g.launch(); torch.cuda.synchronize()
            buf[0] = 0
            torch.cuda.synchronize()
            g.launch()
    else:
        eng._decode_step(st)
    torch.cuda.synchronize()
    ops.debug_set_trace(None)
    n = int(buf[0].item())
    rec = buf[2:2 + 3 * min(n, cap)].view(-1, 3).cpu()
    tag, t = rec[:, 0], rec[:, 1]
    kid, point, cta = (tag >> 48) & 0xFFFF, (tag >> 40) & 0xFF, tag & ((1 << 40) - 1)
    # kernels in launch order: a new launch of the same kernel id begins when CTA 0 reports "start" again
    order = torch.argsort(t, stable=True)
    launches = []           # list of dict(kid, points -> [times])
    open_by_kid = {}
    for i in order.tolist():
        k, p, c, ts = int(kid[i]), int(point[i]), int(cta[i]), int(t[i])
        cur = open_by_kid.get(k)
        if p == 0 and (cur is None or c in cur["seen0"]):
            cur = {"kid": k, "pts": {}, "seen0": set()}
            open_by_kid[k] = cur
            launches.append(cur)
        if cur is None:
            continue
        if p == 0:
            cur["seen0"].add(c)
        cur["pts"].setdefault(p, []).append(ts)
    t0 = min(min(v) for L in launches for v in L["pts"].values())
    rows = []
    for L in launches:
        pts = L["pts"]
        f = lambda p, fn: (fn(pts[p]) - t0) / 1e3 if p in pts else None
        rows.append(dict(kernel=KID.get(L["kid"], str(L["kid"])), ctas=len(L["seen0"]), start=f(0, min), start_last=f(0, max), dep_released=f(1, min),
                         dep_released_last=f(1, max), prologue_done=f(5, max), first_stage=f(2, min), first_stage_last=f(2, max),
                         last_consumed=f(3, max), end=max(x for x in (f(4, max), f(7, max)) if x is not None) if (4 in pts or 7 in pts) else None,
                         rendezvous=f(6, max)))
    step_us = max(r["end"] for r in rows if r["end"] is not None)
    per_layer = [r for r in rows]
    print(json.dumps({"mode": a.mode, "graph": a.graph, "records": n, "kernels_traced": len(rows), "traced_span_us": round(step_us, 1)}))
    # print a window of kernels from the middle of the step
    kpl = 5 if a.mode == "fused" else 7
    lo = 2 * kpl
    hdr = ["kernel", "ctas", "start", "start_last", "dep_released", "dep_released_last", "prologue_done", "first_stage", "first_stage_last", "last_consumed", "rendezvous", "end"]
    print(" | ".join(hdr))
    base = rows[lo]["start"] if len(rows) > lo else 0.0
    for r in rows[lo: lo + a.show * kpl + 1]:
        print(" | ".join(str(r[h]) if h in ("kernel", "ctas") else ("-" if r[h] is None else f"{r[h] - base:.1f}") for h in hdr))
    # aggregate per kernel type: mean of (end - dep_released), (first_stage - dep_released), (last_consumed - first_stage), (end - last_consumed)
    agg = {}
    for r in rows[kpl:]:
        d = agg.setdefault(r["kernel"], {"n": 0, "wall": 0.0, "dep_to_first": 0.0, "stream": 0.0, "tail": 0.0, "start_to_dep": 0.0})
        if None in (r["end"], r["dep_released"], r["first_stage"], r["last_consumed"]):
            continue
        d["n"] += 1
        d["wall"] += r["end"] - r["dep_released"]
        d["dep_to_first"] += r["first_stage_last"] - r["dep_released"]
        d["stream"] += r["last_consumed"] - r["first_stage_last"]
        d["tail"] += r["end"] - r["last_consumed"]
        d["start_to_dep"] += r["dep_released"] - r["start"]
    out = {k: {m: round(v[m] / max(1, v["n"]), 2) for m in ("wall", "dep_to_first", "stream", "tail", "start_to_dep")} | {"n": v["n"]} for k, v in agg.items()}
    print(json.dumps({"per_kernel_mean_us": out}, indent=1))


if __name__ == "__main__":
    main()
