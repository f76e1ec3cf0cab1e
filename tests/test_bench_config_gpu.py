"""Parity ON THE BENCHMARKED CONFIGURATION (BASELINE.json configs[2]: 64 pages of 1024x1024, 512 new tokens, real model size,
`random` N(0, 0.02) weights -- the logits depend on the ViT output, the prefill attention and every KV-cache entry, unlike the
`peaked` checkpoint whose successor is a function of the last token alone).

1. free running, exactly bench.py's call (`make_workload`, CUDA-graph decode): the engine's greedy ids against the reference path's
   (restated ViT + HF Qwen2ForCausalLM.generate in bf16 on the same GPU, parser.py:99-116).  Two correct bf16 pipelines part
   ways at the first near-tie of 152 k Gaussian logits, so the statement is: at each row's FIRST divergence the fp32 oracle's margin
   between the two chosen tokens is below twice the bf16 noise measured at that very position.  The ids' checksum is the one
   bench.py prints (tests/golden/bench_ids_checksum.json pins it).
2. teacher forced on the reference's ids, ragged left-padded prompts (lengths crossing the 64-key tile and the 2048 boundary):
   pre-sampling logits at steps {0, 1, 63, 64, 255, 511} against the fp32 oracle, no worse than 1.5 x HF-bf16's own error;
   argmax equal wherever the fp32 margin clears twice the noise; each checked row also run ALONE (batch 1: other tile shapes,
   other split plans) must give the same logits to bf16 noise.
3. the checks can fail: with the test-only fault switch (decode attention drops the P*V term of every other key tile) criterion 2
   is violated by a wide margin.
"""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = [0, 1, 63, 64, 255, 511]
N_NEW = 512
B = 64
FLOOR = 6e-2            # sigma units: quantisation of bf16 logits alone (see tests/test_engine_gpu.py)


@pytest.fixture(scope="module")
def ctx():
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    from oracle.model import DotsOracle
    cfg = config.full()
    ck = weights.make_synthetic_checkpoint(cfg, 0, "random", device=DEV)
    eng = Engine(cfg, ck, DEV)
    orc16 = DotsOracle(cfg, ck, torch.bfloat16, DEV)
    d = dict(cfg=cfg, ck=ck, eng=eng, orc16=orc16, orc32=None)
    yield d
    d.clear()
    torch.cuda.empty_cache()


def _orc32(ctx):
    if ctx["orc32"] is None:
        from oracle.model import DotsOracle
        ctx["orc32"] = DotsOracle(ctx["cfg"], ctx["ck"], torch.float32, DEV)
    return ctx["orc32"]


@torch.no_grad()
def _tf_logits(orc, emb_prompt, new_row, steps):
    """fp32 logits [len(steps), V] that predict new_row[j], j in steps, for ONE unpadded sequence: prompt embeddings
    [1, T, H] followed by the embeddings of new_row[:max(steps)]."""
    T = emb_prompt.shape[1]
    n = max(steps)
    emb = emb_prompt
    if n > 0:
        emb = torch.cat([emb_prompt, orc.llm.model.embed_tokens(new_row[:n].to(orc.device)).unsqueeze(0).to(emb_prompt.dtype)], dim=1)
    hid = orc.llm.model(inputs_embeds=emb).last_hidden_state[0]
    idx = torch.tensor([T - 1 + j for j in steps], device=hid.device)
    return orc.llm.lm_head(hid[idx]).float()


@torch.no_grad()
def _row_embeds(orc, ids_row, pv_row, grid_row):
    return orc.inputs_embeds(ids_row.unsqueeze(0), pv_row, grid_row)


def test_bench_workload_free_running_ids(ctx):
    sys.path.insert(0, ROOT)
    import bench
    cfg, eng, orc16 = ctx["cfg"], ctx["eng"], ctx["orc16"]
    pv, grid, ids = bench.make_workload(cfg, B, 0, torch.device(DEV))
    T = ids.shape[1]
    s_vit = pv.shape[0] // B
    out = eng.generate(ids.to(DEV), pixel_values=pv, image_grid_thw=grid, max_new_tokens=N_NEW)          # bench.py's step_device()
    got = out.sequences[:, T:].cpu()
    assert got.shape == (B, N_NEW)
    checksum = bench.ids_checksum(got)
    again = eng.generate(ids.to(DEV), pixel_values=pv, image_grid_thw=grid, max_new_tokens=N_NEW).sequences[:, T:].cpu()
    assert torch.equal(got, again), "the engine is not deterministic run to run (cached graph / workspaces)"
    gpath = os.path.join(ROOT, "tests", "golden", "bench_ids_checksum.json")
    key = f"b{B}_n{N_NEW}_p1024_rank0"
    golden = json.load(open(gpath)).get(key) if os.path.exists(gpath) else None
    print(f"bench ids checksum {checksum} (golden {golden})")
    if golden is not None:
        assert checksum == golden, "ids of the benchmark workload changed: re-verify, then update tests/golden/bench_ids_checksum.json"
    # reference path, free running, bf16 on the same GPU
    ref = orc16.generate(ids.to(DEV), pixel_values=pv, image_grid_thw=grid, max_new_tokens=N_NEW)[:, T:].cpu()
    first = []
    for r in range(B):
        ne = (got[r] != ref[r]).nonzero()
        first.append(int(ne[0]) if ne.numel() else N_NEW)
    print("first divergence step per row: min %d median %d max %d; rows identical for all %d steps: %d" % (
        min(first), sorted(first)[B // 2], max(first), N_NEW, sum(f == N_NEW for f in first)))
    # the rows that diverge earliest are the most informative; check 8 of them at their first divergence
    rows = sorted(range(B), key=lambda r: first[r])[:8]
    orc32 = _orc32(ctx)
    worst = 0.0
    for r in rows:
        s = first[r]
        if s == N_NEW:
            continue
        pv_r, grid_r = pv[r * s_vit:(r + 1) * s_vit], grid[r:r + 1]
        l32 = _tf_logits(orc32, _row_embeds(orc32, ids[r].to(DEV), pv_r, grid_r), got[r], [s])[0]
        l16 = _tf_logits(orc16, _row_embeds(orc16, ids[r].to(DEV), pv_r, grid_r), got[r], [s])[0]
        sd = float(l32.std())
        noise = max(float((l16 - l32).abs().max()) / sd, FLOOR)
        margin = abs(float(l32[got[r, s]] - l32[ref[r, s]])) / sd
        top = float(l32.max())
        gap_e, gap_h = (top - float(l32[got[r, s]])) / sd, (top - float(l32[ref[r, s]])) / sd
        print(f"row {r}: diverges at step {s}: fp32 margin between the two tokens {margin:.4f} sigma, bf16 noise there {noise:.4f}; "
              f"distance from the fp32 top-1: engine {gap_e:.4f}, HF-bf16 {gap_h:.4f}")
        assert margin < 2 * noise, (r, s, margin, noise)
        assert gap_e < 2 * noise, (r, s, gap_e, noise)          # the engine's token is itself a near-top candidate of the fp32 oracle
        worst = max(worst, margin / noise)
    print(f"largest margin / noise at a first divergence: {worst:.3f} (must stay below 2)")


def _ragged_inputs(cfg, gen):
    """64 pages of 1024x1024 with text lengths 20..760: prompt lengths 1389..2129, left padded (parser.py:99-105)."""
    from dots_ocr_b200.utils.image_utils import token_counts, vit_grid
    s_vit, t_img = token_counts(1024, 1024, patch=cfg.vision.patch_size, merge=cfg.vision.spatial_merge_size)
    gh, gw = vit_grid(1024, 1024)
    text = [20, 760, 679, 680, 27, 90, 91, 154, 155, 218, 219, 500, 615, 678, 700, 743] + \
        [int(x) for x in torch.randint(20, 761, (B - 16,), generator=gen)]
    rows = []
    for n in text:
        txt = torch.randint(0, 151643, (n,), generator=gen)
        rows.append(torch.cat([txt[: n // 2], torch.full((t_img,), cfg.image_token_id), txt[n // 2:]]))
    Tp = max(r.numel() for r in rows)
    ids = torch.zeros((B, Tp), dtype=torch.long)
    mask = torch.zeros((B, Tp), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, Tp - r.numel():] = r
        mask[i, Tp - r.numel():] = 1
    pv = torch.randn((B * s_vit, cfg.vision.patch_dim), generator=gen)
    grid = torch.tensor([[1, gh, gw]] * B)
    return ids, mask, pv, grid, rows, s_vit


def test_ragged_batch_teacher_forced_logits(ctx):
    from dots_ocr_b200 import ops
    cfg, eng, orc16 = ctx["cfg"], ctx["eng"], ctx["orc16"]
    gen = torch.Generator().manual_seed(2024)
    ids, mask, pv, grid, rows, s_vit = _ragged_inputs(cfg, gen)
    Tp = ids.shape[1]
    lens = mask.sum(1)
    assert int(lens.min()) < 1408 and int(lens.max()) > 2048 and int((lens + N_NEW > 2048).sum()) > 8
    pv_d = pv.to(DEV)
    ref = orc16.generate(ids.to(DEV), attention_mask=mask.to(DEV), pixel_values=pv_d, image_grid_thw=grid, max_new_tokens=N_NEW)
    new = ref[:, Tp:].cpu()
    out = eng.generate(ids.to(DEV), attention_mask=mask.to(DEV), pixel_values=pv_d, image_grid_thw=grid, max_new_tokens=N_NEW,
                       forced_ids=new, return_logits=True)
    assert torch.equal(out.sequences[:, Tp:].cpu(), new)
    logits = out.logits                                                            # [B, N, V] bf16 on the device
    check_rows = [0, 1, 2, 3, 4, 11, 15, 40]
    orc32 = _orc32(ctx)
    worst_ratio, n_clear = 0.0, 0
    kept = {}
    for r in check_rows:
        ids_r = rows[r].to(DEV)
        pv_r, grid_r = pv_d[r * s_vit:(r + 1) * s_vit], grid[r:r + 1]
        l32 = _tf_logits(orc32, _row_embeds(orc32, ids_r, pv_r, grid_r), new[r], STEPS)                 # [6, V]
        l16 = _tf_logits(orc16, _row_embeds(orc16, ids_r, pv_r, grid_r), new[r], STEPS)
        le = logits[r, STEPS].float()
        sd = float(l32.std())
        err_eng = float((le - l32).abs().max()) / sd
        err_hf = float((l16 - l32).abs().max()) / sd
        print(f"row {r} (prompt {rows[r].numel()} tokens): engine-vs-fp32 {err_eng:.4f} sigma, HF-bf16-vs-fp32 {err_hf:.4f}, engine-vs-HF-bf16 "
              f"{float((le - l16).abs().max()) / sd:.4f}")
        assert err_eng < max(1.5 * err_hf, FLOOR), (r, err_eng, err_hf)
        worst_ratio = max(worst_ratio, err_eng / max(err_hf, 1e-9))
        top2 = l32.topk(2, -1).values
        clear = (top2[:, 0] - top2[:, 1]) > 2 * max(err_eng, err_hf) * sd
        n_clear += int(clear.sum())
        assert torch.equal(le.argmax(-1)[clear], l32.argmax(-1)[clear])
        kept[r] = (le.cpu(), sd, err_hf)
    print(f"worst engine/HF error ratio {worst_ratio:.3f}; argmax compared at {n_clear} (row, step) pairs with a clear fp32 margin")
    # the same pages ALONE (batch 1: 32-wide batch tiles, 16 attention splits + combine kernel, other split-K plans)
    for r in check_rows[:4]:
        one = eng.generate(rows[r].unsqueeze(0).to(DEV), pixel_values=pv_d[r * s_vit:(r + 1) * s_vit], image_grid_thw=grid[r:r + 1],
                           max_new_tokens=N_NEW, forced_ids=new[r:r + 1], return_logits=True)
        la = one.logits[0, STEPS].float().cpu()
        le, sd, err_hf = kept[r]
        d = float((la - le).abs().max()) / sd
        print(f"row {r}: alone vs inside the batch of {B}: {d:.4f} sigma")
        assert d < max(1.5 * err_hf, FLOOR), (r, d)
    del logits, out
    # ---- the criterion can fail: fault injection (every other key tile loses its P*V term in decode attention)
    n_short = 66
    try:
        ops.debug_set_fault(2)
        bad = eng.generate(ids.to(DEV), attention_mask=mask.to(DEV), pixel_values=pv_d, image_grid_thw=grid, max_new_tokens=n_short,
                           forced_ids=new[:, :n_short], return_logits=True)
    finally:
        ops.debug_set_fault(0)
    r = check_rows[0]
    l32 = _tf_logits(orc32, _row_embeds(orc32, rows[r].to(DEV), pv_d[r * s_vit:(r + 1) * s_vit], grid[r:r + 1]), new[r], [1, 63, 64])
    sd = float(l32.std())
    err_bad = float((bad.logits[r, [1, 63, 64]].float() - l32).abs().max()) / sd
    print(f"with the injected fault: engine-vs-fp32 {err_bad:.3f} sigma (tolerance {max(1.5 * kept[r][2], FLOOR):.3f})")
    assert err_bad > 3 * max(1.5 * kept[r][2], FLOOR), err_bad
