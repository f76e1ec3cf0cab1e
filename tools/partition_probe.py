#!/usr/bin/env python
"""Go / no-go measurements for running the two phases of the page pipeline side by side on SM partitions (green contexts):

  * ViT encode of --pages pages: whole GPU vs the first partition alone
  * the captured decode step (B = 64, ctx 1881): whole GPU vs the second partition alone
  * both at once, each on its partition (what a 2-deep pipeline would see in steady state)

    python tools/partition_probe.py --first 96 [--pages 16] [--steps 300]
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import config, weights, ops  # noqa: E402
from dots_ocr_b200.engine import Engine  # noqa: E402


def clocks_sampler(stop, out):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.02)
    except Exception as e:      # noqa: BLE001
        out.append(("error", repr(e)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=96, help="SMs of the prefill partition (multiple of 8)")
    ap.add_argument("--pages", type=int, default=16)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1881)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--decode-sms", dest="decode_sms", type=int, default=0, help="plan the decode splits for this many SMs (0: the whole-device plan)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = config.full()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev), dev)
    res = {"first": a.first}
    sp, sd, n_p, n_d = ops.partition(a.first)
    res["sms"] = [n_p, n_d]
    print(json.dumps(res), flush=True)

    gh = gw = 1036 // 14
    S = gh * gw
    pv = torch.randn((a.pages * S, cfg.vision.patch_dim), device=dev)
    grid = [[1, gh, gw]] * a.pages

    def vit():
        return eng.encode_images(pv, grid)

    def time_on(stream, n_sms, fn, reps=1):
        with ops.on_partition(stream, n_sms) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

    ref = vit().clone()
    res["vit_full_ms"] = round(time_on(None, 0, vit), 2)
    print(json.dumps(res), flush=True)
    with ops.on_partition(sp, n_p):
        got = vit()
        torch.cuda.synchronize()
    res["vit_partition_equal"] = bool(torch.equal(ref, got))
    res["vit_first_ms"] = round(time_on(sp, n_p, vit), 2)
    print(json.dumps(res), flush=True)

    # ---- decode graph
    B = a.batch
    ctx_max = (a.ctx + 2 * a.steps + 8 + 63) // 64 * 64
    kc, vc = eng._alloc_cache(B, ctx_max)
    kc.normal_(); vc.normal_()
    lens = torch.full((B,), a.ctx, device=dev, dtype=torch.int64)

    def make_graph(stream, n_sms, decode_sms):
        eng.decode_sms = decode_sms
        st = eng._new_decode_state(B, lens, kc, vc, ctx_max, 2 * a.steps + 8)
        st["last"].copy_(torch.randint(0, 150000, st["last"].shape, generator=torch.Generator().manual_seed(7)).to(st["last"].dtype))
        st["_init"] = {k: st[k].clone() for k in ("step", "pos", "ctx_len", "last")}
        torch.cuda.synchronize()
        ctxm = ops.on_partition(stream, n_sms) if stream is not None else torch.cuda.stream(eng._cap_stream)
        with ctxm:
            eng._decode_step(st)
            torch.cuda.synchronize()
            g = ops.capture(lambda: eng._decode_step(st))
        eng.decode_sms = 0
        return g, st

    def rewind(st):
        for k, v in st["_init"].items():
            st[k].copy_(v)

    def time_graph(g, stream, n):
        with torch.cuda.stream(stream):
            for _ in range(5):
                g.launch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                g.launch()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

    g_full, st_full = make_graph(None, 0, 0)
    res["decode_full_ms"] = round(time_graph(g_full, eng._cap_stream, a.steps), 4)
    print(json.dumps(res), flush=True)
    ids_full = st_full["out_ids"][:, :50].clone()
    g_d, st_d = make_graph(sd, n_d, a.decode_sms)
    res["decode_rest_ms"] = round(time_graph(g_d, sd, a.steps), 4)
    res["decode_rest_plan"] = {k: v for k, v in st_d["plan"].items()}
    print(json.dumps(res), flush=True)
    # same tokens from both graphs?  (same plan -> same arithmetic)
    res["decode_ids_equal"] = bool(torch.equal(ids_full, st_d["out_ids"][:, :50])) if a.decode_sms == 0 else None

    # ---- both at once: decode steps on the rest (enqueued first: a replay costs the host ~10 us), ViT passes on the first partition inside
    # that window; decode is timed in chunks of --steps launches so that the chunks overlapped by ViT work can be told apart
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    vit_reps = max(1, int(1500 / max(res["vit_first_ms"], 1)))
    n_chunks = max(2, int(2200 / max(res["decode_rest_ms"] * a.steps, 1)))
    torch.cuda.synchronize()
    th.start()
    marks = []
    with torch.cuda.stream(sd):
        for c in range(n_chunks):
            rewind(st_d)                        # keep the context length of the measured steps at --ctx .. --ctx + --steps
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            for _ in range(a.steps):
                g_d.launch()
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ops.on_partition(sp, n_p):
        ev[0].record()
        for _ in range(vit_reps):
            vit()
        ev[1].record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    res["both_vit_ms"] = round(ev[0].elapsed_time(ev[1]) / vit_reps, 2)
    chunks = [round(x.elapsed_time(y) / a.steps, 4) for x, y in zip(marks[:-1], marks[1:])]
    t_v0, t_v1 = marks[0].elapsed_time(ev[0]), marks[0].elapsed_time(ev[1])         # ViT window relative to the first decode chunk
    inside = [c for i, c in enumerate(chunks) if marks[0].elapsed_time(marks[i]) >= t_v0 and marks[0].elapsed_time(marks[i + 1]) <= t_v1]
    res["both_decode_chunks_ms"] = chunks
    res["both_decode_ms"] = sorted(inside)[len(inside) // 2] if inside else None
    res["both_windows_ms"] = {"vit": [round(t_v0, 1), round(t_v1, 1)], "decode_end": round(marks[0].elapsed_time(marks[-1]), 1)}
    good = [s for s in samples if s[0] != "error"]
    if good:
        mid = good[len(good) // 4: 3 * len(good) // 4] or good
        res["both_sm_mhz_median"] = sorted(s[0] for s in mid)[len(mid) // 2]
        res["both_power_w_median"] = round(sorted(s[1] for s in mid)[len(mid) // 2], 1)
    else:
        res["clock_sampler"] = samples[:1]
    # projected steady state of a 2-deep pipeline at B = 64: per batch max(prefill on the first partition, 511 decode steps on the rest)
    per_page_vit = res["both_vit_ms"] / a.pages
    res["projection"] = {"vit_64_pages_ms": round(per_page_vit * 64, 1), "decode_511_steps_ms": round((res["both_decode_ms"] or 0) * 511, 1)}
    print(json.dumps(res), flush=True)
    del g_d, g_full
    torch.cuda.synchronize()
    ops.partition_destroy()


if __name__ == "__main__":
    main()
