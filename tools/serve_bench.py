#!/usr/bin/env python
"""Serving-path throughput on a MIXED-LENGTH workload: the reference parser's fan-out -- many threads, one page
per ``inference_with_vllm`` call, every page with its own output length (parser.py:282-290) -- through

  * ``BatchingRunner``      batches formed at arrival, every batch runs to its LONGEST page before the queue is looked at again
  * ``ContinuousBatcher``   fixed decode slots, finished rows are harvested and refilled between chunks of decode steps

Synthetic weights never emit a stop id, so a page's length is its token budget: budgets are drawn uniformly from
[--min-new, --max-new] (seeded), pages are 1024x1024 uint8.  Reports pages/s and generated tokens/s of each arm and checks
that both give every page the same ids.

    python tools/serve_bench.py --pages 192 --min-new 100 --max-new 2000 --threads 64
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=192)
    ap.add_argument("--min-new", dest="min_new", type=int, default=100)
    ap.add_argument("--max-new", dest="max_new", type=int, default=2000)
    ap.add_argument("--threads", type=int, default=64)
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--preset", default="full")
    ap.add_argument("--arms", default="batching,continuous")
    a = ap.parse_args()
    from PIL import Image
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.batching import BatchingRunner
    from dots_ocr_b200.continuous import ContinuousBatcher, EngineSlots
    from dots_ocr_b200.engine import Engine
    from dots_ocr_b200.processing import SyntheticTokenizer
    from dots_ocr_b200.runner import PageRunner

    class IdTokenizer(SyntheticTokenizer):           # "text" = the ids, so that the two arms can be compared exactly
        def decode(self, ids):
            return ",".join(str(int(i)) for i in ids)

    dev = torch.device("cuda:0")
    cfg = config.PRESETS[a.preset]()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev), dev)
    tok = IdTokenizer(cfg)
    rng = np.random.default_rng(0)
    pages = [Image.fromarray(rng.integers(0, 256, (a.side, a.side, 3), dtype=np.uint8)) for _ in range(min(a.pages, 16))]
    budgets = [int(x) for x in rng.integers(a.min_new, a.max_new + 1, a.pages)]
    prompt = "Please output the layout information from the PDF image."
    results = {}
    outs = {}
    for arm in a.arms.split(","):
        if arm == "batching":
            front = BatchingRunner(PageRunner(eng, tok), max_batch=a.slots, max_wait_ms=20)
        else:
            front = ContinuousBatcher(EngineSlots(eng, tok, n_slots=a.slots, max_prompt=2048, max_new=a.max_new, chunk=a.chunk))
        got = [None] * a.pages
        nxt = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= a.pages:
                    return
                got[i] = front.infer(pages[i % len(pages)], prompt, max_new_tokens=budgets[i])

        # warm-up: kernels, graph, allocator
        front.infer(pages[0], prompt, max_new_tokens=8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts = [threading.Thread(target=worker) for _ in range(a.threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        front.close()
        toks = sum(budgets)
        results[arm] = {"seconds": round(dt, 3), "pages_per_s": round(a.pages / dt, 3), "tokens_per_s": round(toks / dt, 1)}
        if arm == "continuous":
            results[arm]["stats"] = dict(front.stats)
        else:
            results[arm]["batches"] = len(front.batches)
        outs[arm] = got
        assert all(g is not None and len(g.split(",")) == b for g, b in zip(got, budgets)), f"{arm}: a page came back with the wrong length"
    line = {"workload": f"{a.pages} pages {a.side}x{a.side}, budgets U[{a.min_new},{a.max_new}] tokens (mean {sum(budgets) / len(budgets):.0f}), "
                        f"{a.threads} caller threads, {a.slots} slots / max batch", "results": results}
    if len(outs) == 2:
        x, y = outs["batching"], outs["continuous"]
        # random weights: two correct schedules agree until a near-tie flips a token; report how far the rows agree
        same = sum(1 for p, q in zip(x, y) if p == q)
        line["rows_identical"] = same
        line["speedup_continuous_over_batching"] = round(results["continuous"]["pages_per_s"] / results["batching"]["pages_per_s"], 3)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
