#!/usr/bin/env python
"""Run a few eager decode steps at the benchmark's decode geometry (batch 64, ~1.9 k keys of context, full model) so that ncu can
list every kernel of the step with its duration and DRAM traffic:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dots:: \
        --csv --log-file gpurun_out/decode_step.csv python tools/decode_step_profile.py --steps 3
    python tools/decode_step_profile.py --summarise gpurun_out/decode_step.csv --steps 3 > profiles/decode_traffic_rNN.json
"""
import argparse
import csv
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path: str, steps: int) -> dict:
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = {}
    for r in rows:
        name = r.get("Kernel Name", "")
        if "dots::" not in name and not name.startswith("dots"):
            continue
        k = name.split("(")[0].replace("void ", "").strip()
        d = per.setdefault(k, {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["launches"] += 1
            d["time_us"] += val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3 if unit in ("ms", "msecond") else val)
        elif m == "dram__bytes_read.sum":
            d["dram_read"] += val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif m == "dram__bytes_write.sum":
            d["dram_write"] += val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    out = {"steps": steps, "kernels": {}}
    tot_t = tot_b = 0.0
    for k, d in per.items():
        out["kernels"][k] = {"launches_per_step": d["launches"] / steps, "avg_us": round(d["time_us"] / max(1, d["launches"]), 2),
                             "us_per_step": round(d["time_us"] / steps, 1), "dram_MB_per_step": round((d["dram_read"] + d["dram_write"]) / steps / 1e6, 2)}
        tot_t += d["time_us"] / steps
        tot_b += (d["dram_read"] + d["dram_write"]) / steps
    out["serialised_us_per_step"] = round(tot_t, 1)
    out["dram_bytes_per_step"] = round(tot_b)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1881)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default=None)
    ap.add_argument("--attn-splits", dest="attn_splits", type=int, default=0)
    ap.add_argument("--summarise", default=None)
    a = ap.parse_args()
    if a.summarise:
        print(json.dumps(summarise(a.summarise, a.steps), indent=1))
        return
    import torch
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    dev = torch.device("cuda:0")
    cfg = config.full()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev), dev)
    if a.mode:
        eng.decode_mode = a.mode
    eng.attn_splits = a.attn_splits
    ctx_max = (a.ctx + a.steps + 2 + 63) // 64 * 64
    kc, vc = eng._alloc_cache(a.batch, ctx_max)
    kc.normal_(); vc.normal_()
    lens = torch.full((a.batch,), a.ctx, device=dev, dtype=torch.int64)
    st = eng._new_decode_state(a.batch, lens, kc, vc, ctx_max, a.steps + 2)
    st["last"].random_(0, 150000)
    torch.cuda.synchronize()
    for _ in range(a.steps):
        eng._decode_step(st)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
