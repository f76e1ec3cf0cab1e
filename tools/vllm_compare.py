#!/usr/bin/env python
"""On-box comparison against the in-image vLLM ``DotsOCRForCausalLM`` on the SAME synthetic parameters.

Results: profiles/vllm_compare_r2.md.

    python tools/make_checkpoint_dir.py --preset full --flavour peaked --out /tmp/dots_full
    python tools/vllm_compare.py --dir /tmp/dots_full --impl vllm --pages 64 --new-tokens 512 > gpurun_out/vllm.json
    python tools/vllm_compare.py --dir /tmp/dots_full --impl ours --pages 64 --new-tokens 512 > gpurun_out/ours.json
    python tools/vllm_compare.py --diff gpurun_out/vllm.json gpurun_out/ours.json

Both arms take the same seeded uint8 pages as PIL images and the same prompt, decode greedily for exactly --new-tokens
tokens (stop ids ignored) and are timed by wall clock around the whole call (image processor included) after one warm-up
call; each writes one JSON line with pages/s and the generated ids.  ``--diff`` reports the fraction of identical ids (with
the `peaked` flavour the successor of every token is baked into lm_head, so the ids must agree exactly; with `random` they
agree until the first near-tie, since the two image processors differ in the last bit of the bicubic resize).

The two arms are separate processes so that each owns the GPU alone.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PROMPT = "Please output the layout information from the PDF image, including each layout element's bbox, its category, " \
         "and the corresponding text content within the bbox."


def pages(n: int, side: int, seed: int = 0):
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    return [Image.fromarray(rng.integers(0, 256, (side, side, 3), dtype=np.uint8)) for _ in range(n)]


def run_vllm(a) -> dict:
    from vllm import LLM, SamplingParams
    from dots_ocr_b200.processing import HFTokenizer
    text = HFTokenizer(a.dir).render(PROMPT)                     # ...<|img|><|imgpad|><|endofimg|>{prompt}...; vLLM widens the pad
    llm = LLM(model=a.dir, tokenizer=a.dir, dtype="bfloat16", trust_remote_code=True, max_model_len=a.max_model_len, max_num_seqs=a.pages,
              limit_mm_per_prompt={"image": 1}, gpu_memory_utilization=a.gpu_mem, enable_prefix_caching=False,
              enforce_eager=a.eager)
    sp = SamplingParams(temperature=0.0, max_tokens=a.new_tokens, ignore_eos=True, detokenize=False)
    imgs = pages(a.pages, a.side)
    reqs = [{"prompt": text, "multi_modal_data": {"image": im}} for im in imgs]
    llm.generate(reqs[: max(1, min(4, a.pages))], SamplingParams(temperature=0.0, max_tokens=8, ignore_eos=True, detokenize=False))
    best, ids = None, None
    for _ in range(a.repeats):
        t0 = time.perf_counter()
        outs = llm.generate(reqs, sp)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        ids = [list(o.outputs[0].token_ids) for o in outs]
    return {"impl": "vllm", "seconds": best, "ids": ids}


def run_ours(a) -> dict:
    import torch
    from dots_ocr_b200.processing import build_text_inputs, page_to_u8, model_image_tokens
    from dots_ocr_b200.runner import PageRunner
    runner = PageRunner.from_checkpoint(a.dir, device="cuda:0")
    eng, tk = runner.engine, runner.tokenizer
    imgs = pages(a.pages, a.side)

    def once(batch, n_new):
        pg = [page_to_u8(im) for im in batch]                 # RGB bytes at the original size; resize + normalise + patchify on the GPU
        inp = build_text_inputs(tk, [model_image_tokens(int(p.shape[0]), int(p.shape[1])) for p in pg], [PROMPT] * len(batch))
        out = eng.generate(input_ids=inp["input_ids"].to(eng.device), attention_mask=inp["attention_mask"].to(eng.device),
                           pages_u8=[p.pin_memory().to(eng.device, non_blocking=True) for p in pg], max_new_tokens=n_new,
                           eos_token_id=None, pad_token_id=tk.pad_token_id)
        T = inp["input_ids"].shape[1]
        return out.sequences[:, T:].cpu()

    once(imgs[: max(1, min(4, a.pages))], 8)
    best, ids = None, None
    for _ in range(a.repeats):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        new = once(imgs, a.new_tokens)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        ids = new.tolist()
    return {"impl": "ours", "seconds": best, "ids": ids}


def diff(pa: str, pb: str) -> dict:
    def last_json(p):
        with open(p) as f:
            lines = [ln for ln in f.read().splitlines() if ln.startswith("{")]
        return json.loads(lines[-1])
    A, B = last_json(pa), last_json(pb)
    same = total = first_bad = 0
    firsts = []
    for ra, rb in zip(A["ids"], B["ids"]):
        n = min(len(ra), len(rb))
        eq = [x == y for x, y in zip(ra[:n], rb[:n])]
        same += sum(eq)
        total += n
        firsts.append(eq.index(False) if False in eq else n)
    return {"pages": len(firsts), "ids_equal_frac": same / max(1, total), "first_mismatch_min": min(firsts) if firsts else None,
            A["impl"] + "_pages_per_s": A["pages_per_s"], B["impl"] + "_pages_per_s": B["pages_per_s"],
            "speedup_" + B["impl"] + "_over_" + A["impl"]: B["pages_per_s"] / A["pages_per_s"]}


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--dir", help="checkpoint directory made by tools/make_checkpoint_dir.py")
    ap.add_argument("--impl", choices=("vllm", "ours"))
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--new-tokens", type=int, default=512)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--max-model-len", type=int, default=4096)
    ap.add_argument("--gpu-mem", type=float, default=0.85)
    ap.add_argument("--eager", action="store_true", help="vLLM without its own CUDA graphs / compile step")
    ap.add_argument("--diff", nargs=2, metavar=("A.json", "B.json"))
    a = ap.parse_args()
    if a.diff:
        print(json.dumps(diff(*a.diff)))
        return
    if not a.dir or not a.impl:
        ap.error("--dir and --impl are required unless --diff is given")
    r = run_vllm(a) if a.impl == "vllm" else run_ours(a)
    r.update(pages=a.pages, side=a.side, new_tokens=a.new_tokens, pages_per_s=a.pages / r["seconds"])
    print(json.dumps(r))


if __name__ == "__main__":
    main()
