#!/usr/bin/env python
"""Headline benchmark: pages/sec for the dots.ocr page-parsing hot path on B200.

One "step" = one pass of the whole hot path over one batch of synthetic pages on every rank:
ViT encode (42 blocks, 5476 patch tokens per 1024x1024 page) -> embed + image scatter -> LLM prefill
(T = 1625) -> greedy decode of NEW_TOKENS tokens (EOS disabled: random weights never emit it).
Workload = BASELINE.json configs[2] ("batch=64 synthetic 1024x1024 pages, bf16, greedy, 1xB200");
at N GPUs every rank runs its own 64-page shard (pages are independent: weak scaling, no collective
on the data path; configs[3] = 512 pages on 8 GPUs).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the HF/PyTorch CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGE_HW = (1024, 1024)
TEXT_TOKENS = 256
METRIC = "pages/sec (1024x1024 doc images, greedy)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def _page_tokens(cfg):
    from dots_ocr_b200.utils.image_utils import token_counts
    return token_counts(*PAGE_HW, patch=cfg.vision.patch_size, merge=cfg.vision.spatial_merge_size)


def _prompt_len():
    from dots_ocr_b200.utils.image_utils import token_counts
    return token_counts(*PAGE_HW)[1] + TEXT_TOKENS


def _prompt_ids(cfg, batch, t_img, seed=7):
    import torch
    g = torch.Generator().manual_seed(seed)
    vocab_text = min(cfg.text.vocab_size, 151643)
    front = TEXT_TOKENS // 2
    rows = []
    for _ in range(batch):
        txt = torch.randint(0, vocab_text, (TEXT_TOKENS,), generator=g)
        rows.append(torch.cat([txt[:front], torch.full((t_img,), cfg.image_token_id), txt[front:]]))
    return torch.stack(rows)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def usable_cpu_threads() -> int:
    """Host threads this process may really use: the scheduler affinity mask, capped by a cgroup CPU quota if one is set
    (os.cpu_count() reports the machine, not the container: oversubscribing a quota makes PyTorch's CPU path crawl)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference(new_tokens: int, threads: int, decode_steps: int = 8):
    """The reference's HF CPU float32 path (oracle port: restated ViT + HF Qwen2ForCausalLM) on the host
    cores, on a BOUNDED sample of the same workload: one 1024x1024 page through 2 of 42 ViT blocks and 2
    of 28 decoder layers at full width (plus patch-embed, merger, lm_head), `decode_steps` greedy steps;
    per-layer times are then scaled by the real layer counts."""
    import torch
    from dots_ocr_b200 import config, weights
    from oracle.model import DotsOracle
    torch.set_num_threads(threads)
    full, small = config.full(), config.small()
    s_vit, t_img = _page_tokens(full)
    ck = weights.make_synthetic_checkpoint(small, 0, "random", device="cpu")
    orc = DotsOracle(small, ck, torch.float32, "cpu")
    del ck
    g = torch.Generator().manual_seed(1234)
    pv = torch.randn(s_vit, full.vision.patch_dim, generator=g)
    grid = torch.tensor([[1, PAGE_HW[0] * 1036 // 1024 // 14, PAGE_HW[1] * 1036 // 1024 // 14]])
    from dots_ocr_b200.utils.image_utils import vit_grid
    gh, gw = vit_grid(*PAGE_HW)
    grid = torch.tensor([[1, gh, gw]])
    ids = _prompt_ids(full, 1, t_img)
    v = orc.vision

    def clock(fn):
        t0 = time.perf_counter(); r = fn(); return r, time.perf_counter() - t0

    with torch.no_grad():
        # ViT: fixed part (patch embed + post norm + merger) and per-block part
        from oracle.vision import rot_pos_emb, rms_norm
        ang = rot_pos_emb(grid.tolist(), 2, full.vision.head_dim, full.vision.rope_theta, "cpu")
        cos, sin = ang.cos(), ang.sin()
        x, t_pe = clock(lambda: v.patch_embed(pv))
        cu = [0, s_vit]
        x, t_b0 = clock(lambda: v.block(0, x, cu, cos, sin))
        x, t_b1 = clock(lambda: v.block(1, x, cu, cos, sin))
        img, t_mg = clock(lambda: v.merger(rms_norm(x, v.w["post_trunk_norm.weight"], full.vision.rms_norm_eps)))
        t_vit_layer = min(t_b0, t_b1)
        t_vit = t_pe + t_mg + full.vision.num_hidden_layers * t_vit_layer
        # prefill: embeddings -> 2 layers -> head; per-layer time from the 2-layer model
        emb = orc.llm.model.embed_tokens(ids)
        emb = emb.masked_scatter((ids == full.image_token_id).unsqueeze(-1).expand_as(emb), img.to(emb.dtype))
        from transformers import DynamicCache
        cache = DynamicCache(config=orc.llm.config)
        out, t_pf2 = clock(lambda: orc.llm(inputs_embeds=emb, past_key_values=cache, use_cache=True, logits_to_keep=1))
        _, t_head = clock(lambda: orc.llm.lm_head(out.logits.new_zeros(1, 1, full.text.hidden_size)))
        t_prefill_layer = max(1e-9, (t_pf2 - t_head) / 2)
        t_prefill = full.text.num_hidden_layers * t_prefill_layer + t_head
        # decode steps on the 2-layer model with a KV cache.  One-token steps are tiny memory-bound ops: a very wide thread
        # pool can be slower than a moderate one, so the reference gets the better of {all usable threads, 16 threads}.
        nxt = out.logits[:, -1].argmax(-1, keepdim=True)
        best = None
        for nthr in sorted({threads, min(threads, 16)}, reverse=True):
            torch.set_num_threads(nthr)
            ts = []
            for _ in range(decode_steps):
                o, dt = clock(lambda: orc.llm(input_ids=nxt, past_key_values=cache, use_cache=True))
                nxt = o.logits[:, -1].argmax(-1, keepdim=True)
                ts.append(dt)
            ts.sort()
            med = ts[len(ts) // 2]
            if best is None or med < best[0]:
                best = (med, nthr)
        torch.set_num_threads(threads)
        t_step2, decode_threads = best
        _, t_head1 = clock(lambda: orc.llm.lm_head(out.logits.new_zeros(1, 1, full.text.hidden_size)))
        t_dec_layer = max(1e-9, (t_step2 - min(t_head, t_head1)) / 2)
        t_step = full.text.num_hidden_layers * t_dec_layer + min(t_head, t_head1)
    total = t_vit + t_prefill + new_tokens * t_step
    return dict(pages_per_sec=1.0 / total, t_vit=t_vit, t_prefill=t_prefill, t_step=t_step,
                sample=(f"1 page {PAGE_HW[0]}x{PAGE_HW[1]}: 2/42 ViT blocks + 2/28 decoder layers at full width (fp32, "
                        f"{threads} threads; decode steps with {decode_threads}), {decode_steps} decode steps; per-layer times scaled to 42/28 layers, "
                        f"N={new_tokens} new tokens"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = usable_cpu_threads()
    vals, t0 = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        r = cpu_reference(args.new_tokens, threads, decode_steps=4)
        if i >= args.warmup:
            vals.append(r)
        if time.perf_counter() - t0 > 240 and vals:
            break
    v = sum(x["pages_per_sec"] for x in vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pages/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _config(args, 1),
            "cpu_baseline": {"value": v, "unit": "pages/s", "cores": threads, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _config(args, world):
    return {"workload": f"batch={args.batch} synthetic {PAGE_HW[0]}x{PAGE_HW[1]} pages per GPU, ViT encode + prefill(T={_prompt_len()}) + "
                        f"greedy decode N={args.new_tokens} (BASELINE configs[2]; x{world} GPUs = configs[3] sharding)",
            "pages_per_gpu": args.batch, "global_pages": args.batch * world, "new_tokens": args.new_tokens,
            "prompt_tokens": _prompt_len(), "parallelism": f"dp{world} (page shards, model replicated, no collective)",
            "l2": "inputs larger than L2 (pixel_values 824 MB fp32 per step; weights 6.1 GB streamed every decode step)",
            "weights": "synthetic N(0,0.02) seed 0, real architecture (ViT 1.26B + LLM 1.78B)"}


# --------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line: NCCL_DEBUG=VERSION/INFO in the environment prints a banner on stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    from dots_ocr_b200 import config, weights, ops
    from dots_ocr_b200.engine import Engine
    if args.attn_impl:
        ops.ATTN_IMPL = args.attn_impl
    if args.no_pdl:
        ops.set_pdl(False)
    if args.gemm_pair is not None:
        ops.set_gemm_pair(bool(args.gemm_pair))
    cfg = config.PRESETS[args.preset]()
    ck = weights.make_synthetic_checkpoint(cfg, 0, "random", device=dev)
    eng = Engine(cfg, ck, dev)
    del ck
    torch.cuda.empty_cache()

    s_vit, t_img = _page_tokens(cfg)
    from dots_ocr_b200.utils.image_utils import vit_grid
    gh, gw = vit_grid(*PAGE_HW)
    B, N = args.batch, args.new_tokens
    grid = torch.tensor([[1, gh, gw]] * B)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pv_dev = torch.randn((B * s_vit, cfg.vision.patch_dim), generator=g, device=dev)     # normalised pixel_values ~ N(0,1)
    ids = _prompt_ids(cfg, B, t_img)
    ids_dev = ids.to(dev)
    pv_host = torch.empty(pv_dev.shape, dtype=torch.float32, pin_memory=True)
    pv_host.copy_(pv_dev)
    ids_host = ids.pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return eng.generate(ids_dev, pixel_values=pv_dev, image_grid_thw=grid, max_new_tokens=N)

    def step_e2e():
        pv = pv_host.to(dev, non_blocking=True)
        idd = ids_host.to(dev, non_blocking=True)
        out = eng.generate(idd, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N)
        return out.sequences[:, idd.shape[1]:].cpu()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    # ---- timed region 1: inputs resident in HBM; per-kernel CUDA events via the ops profiling hook
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ops.PROFILE = prof = []
    ms_total = timed(step_device, args.steps)
    ops.PROFILE = None
    launches = eng.launches - l0
    # ---- timed region 2: end to end through the public API with pinned host buffers
    ms_e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    pages = B * world * args.steps
    value = pages / (ms_total / 1e3)
    e2e_v = pages / (ms_e2e / 1e3) if ms_e2e else None
    h2d = pv_host.numel() * 4 + ids_host.numel() * 8
    d2h = B * N * 8

    # ---- roofline of the dominant kernel (tcgen05 GEMM, all prefill-side shapes) from the live events
    peaks = _peaks()
    torch.cuda.synchronize()
    agg = {}
    for name, work, e0, e1 in prof:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += work; a[1] += e0.elapsed_time(e1); a[2] += 1
    roof = None
    by_kernel = {}
    for name, (work, ms, n) in agg.items():
        by_kernel[name] = {"launches": n, "ms": round(ms, 3), "share_of_step": round(ms / ms_total, 4)}
    if "gemm_bf16_tcgen05" in agg:
        fl, ms, n = agg["gemm_bf16_tcgen05"]
        ach = fl / (ms / 1e3) / 1e12
        roof = {"kernel": "gemm_bf16_tcgen05_kernel (ViT + LLM prefill linears)", "bound": "tensor", "achieved": round(ach, 1),
                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sustained"], 4),
                "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; burst {peaks['tf_burst']})", "traffic": None,
                "launches": n, "avg_launch_ms": round(ms / n, 4)}
        by_kernel["gemm_bf16_tcgen05"]["tflops"] = round(ach, 1)
    if "attn_fwd_vit" in agg:
        fl, ms, n = agg["attn_fwd_vit"]
        by_kernel["attn_fwd_vit"]["tflops"] = round(fl / (ms / 1e3) / 1e12, 1)
    roof_dec = None
    if "decode_phase" in agg:
        by, ms, n = agg["decode_phase"]           # bytes of all decode steps, ms of all decode phases
        ach = by / (ms / 1e3) / 1e9
        roof_dec = {"kernel": "decode step (CUDA graph: skinny tcgen05 GEMMs + KV attention + finalize kernels)", "bound": "hbm",
                    "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4),
                    "ms_per_decode_step": round(ms / (n * max(1, N - 1)), 4), "traffic": None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    threads = usable_cpu_threads()
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference(N, threads, decode_steps=8)
            cpu = {"value": r["pages_per_sec"], "unit": "pages/s", "cores": threads, "kind": "port", "sample": r["sample"],
                   "t_vit_s": round(r["t_vit"], 2), "t_prefill_s": round(r["t_prefill"], 2), "t_step_s": round(r["t_step"], 4)}
        except Exception as e:      # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "pages/s", "cores": threads, "kind": "port", "sample": f"failed: {e!r}"}
    line = {"metric": METRIC, "value": round(value, 3), "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": _config(args, world),
            "e2e": {"value": round(e2e_v, 3) if e2e_v else None, "unit": "pages/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_decode": roof_dec, "kernels": by_kernel,
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--new-tokens", dest="new_tokens", type=int, default=512)
    ap.add_argument("--preset", default="full")
    ap.add_argument("--page", type=int, default=1024, help="synthetic page edge in pixels (1024 = the headline workload; 1960 with "
                    "--batch 4 --new-tokens 2048 = BASELINE configs[4], the long-context case)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--gemm-pair", dest="gemm_pair", type=int, default=None, choices=[0, 1],
                    help="CTA-pair (cta_group::2) kernel for the large prefill GEMMs (default: the library default)")
    ap.add_argument("--no-pdl", dest="no_pdl", action="store_true", help="plain stream order between kernels (A/B runs)")
    ap.add_argument("--attn-impl", dest="attn_impl", default=None, choices=["tc", "mma"])
    args = ap.parse_args()
    global PAGE_HW
    PAGE_HW = (args.page, args.page)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
