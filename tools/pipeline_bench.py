#!/usr/bin/env python
"""The benchmark workload (64 synthetic 1024 x 1024 pages per batch, 512 new tokens) through
  * Engine.generate, batch after batch (what bench.py times), and
  * PagePipeline: batch i+1's encode + prefill on one SM partition while batch i decodes on the other,
K batches each, bracketed by synchronize; ids of every batch compared.  One JSON line.

    python tools/pipeline_bench.py --first 96 [--batches 4] [--batch 64] [--new-tokens 512]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dots_ocr_b200 import config, weights  # noqa: E402
from dots_ocr_b200.engine import Engine  # noqa: E402
from dots_ocr_b200.pipeline import PagePipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=96)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--new-tokens", dest="new_tokens", type=int, default=512)
    ap.add_argument("--decode-plan-sms", dest="plan", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = config.full()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev), dev)
    pv, grid, ids = bench.make_workload(cfg, a.batch, 0, dev)
    req = dict(input_ids=ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=a.new_tokens)
    K = a.batches

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1)

    seq_one = lambda: [eng.generate(ids.to(dev), pixel_values=pv, image_grid_thw=grid, max_new_tokens=a.new_tokens).sequences for _ in range(K)]
    seq_one()
    want, ms_seq = timed(seq_one)
    res = {"batches": K, "pages_per_batch": a.batch, "new_tokens": a.new_tokens,
           "sequential": {"ms": round(ms_seq, 1), "pages_per_s": round(K * a.batch / ms_seq * 1e3, 3)}}
    print(json.dumps(res), flush=True)
    with PagePipeline(eng, prefill_sms=a.first, decode_plan_sms=a.plan) as pipe:
        res["sms"] = [pipe.n_pre, pipe.n_dec]
        pipe.run([req] * 3)                                   # slots, graphs, allocator pools of both partition streams
        eng.decode_log.clear()
        got, ms_pipe = timed(lambda: pipe.run([req] * K))
        dec = [round(e0.elapsed_time(e1) / max(1, steps), 4) for (_b, steps, e0, e1) in eng.decode_log[-K:]]
    res["pipelined"] = {"ms": round(ms_pipe, 1), "pages_per_s": round(K * a.batch / ms_pipe * 1e3, 3), "decode_ms_per_step": dec}
    res["ids_equal"] = all(torch.equal(w, g.sequences) for w, g in zip(want, got))
    res["speedup"] = round(ms_seq / ms_pipe, 3)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
