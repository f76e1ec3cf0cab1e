"""The 5-kernel decode layer (cluster split-K GEMMs with on-chip reduction + residual + RMSNorm, cluster-merged attention)
against (a) a plain PyTorch restatement with HF's rounding points and (b) the per-op kernels it replaces.

Bit-exactness statements: with the same K partition (8 splits) the cluster GEMM adds the same fp32 partials in the same
order as dots_gemm_skinny_bf16 + its finalize kernel, so q|k|v, the new residual stream and the attention output are
bit-identical; only the RMSNorm statistic is summed in a different order (per 128-feature tile), which may move `normed`
by one bf16 ulp on a small fraction of elements."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from dots_ocr_b200 import ops
    return ops


def _bf(x):
    return x.to(torch.bfloat16)


def _rand(shape, gen, scale=1.0):
    return _bf(torch.randn(shape, generator=gen, device=DEV) * scale)


@pytest.fixture(scope="module")
def gen():
    g = torch.Generator(device=DEV)
    g.manual_seed(4321)
    return g


def _hf_rope_bf16(x, pos, inv_freq):
    # [Q] modeling_qwen2.py:102-146: cos/sin in bf16, q*cos + rotate_half(q)*sin in bf16
    freqs = pos.float()[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(torch.bfloat16)[:, None, :], emb.sin().to(torch.bfloat16)[:, None, :]
    x1, x2 = x[..., :64], x[..., 64:]
    rot = torch.cat((-x2, x1), dim=-1)
    return (x * cos) + (rot * sin)


def _ulp_diff(a, b):
    """bf16 tensors -> |difference| in units of bf16 steps of the larger magnitude (0 = equal)."""
    ai = a.view(torch.int16).to(torch.int32)
    bi = b.view(torch.int16).to(torch.int32)
    return (ai - bi).abs()


@pytest.mark.parametrize("B,N,K", [(64, 2048, 1536), (33, 2048, 1536), (1, 2048, 1536), (7, 1024, 768), (64, 256, 128), (20, 2048, 1536)])
def test_decode_gemm_qkv(B, N, K, gen):
    ops = _ops()
    x, w, bias = _rand((B, K), gen), _rand((N, K), gen, 0.03), _rand((N,), gen)
    out = torch.full((B, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    R = ops.decode_tile_rows(B)
    ops.decode_gemm_qkv(ops.tile_rows(x, R).view(-1), ops.tile_weight(w), bias, out, K)
    ref = _bf(x.float() @ w.float().t() + bias.float())
    assert not torch.isnan(out.float()).any()
    err = float((out.float() - ref.float()).abs().max() / ref.float().abs().max())
    assert err < 8e-3, err
    assert float((out == ref).float().mean()) > 0.97
    # the per-op path with the same K partition: identical partial sums, identical order
    nkb = -(-K // 64)
    per = -(-nkb // 8)
    splits = -(-nkb // per)
    part = ops.gemm_skinny(x, w, splits)
    acc = part[0].clone()
    for s in range(1, splits):
        acc += part[s]
    old = _bf(acc + bias.float())
    assert torch.equal(out, old)


@pytest.mark.parametrize("B,N,K", [(64, 1536, 1536), (64, 1536, 8960), (33, 1536, 4224), (1, 1536, 8960), (7, 768, 1024), (20, 768, 768)])
def test_decode_gemm_resnorm(B, N, K, gen):
    ops = _ops()
    eps = 1e-6
    x, w = _rand((B, K), gen), _rand((N, K), gen, 0.03)
    resid0 = _rand((B, N), gen)
    ln_w = _bf(1 + 0.1 * torch.randn(N, generator=gen, device=DEV))
    tiles = -(-N // 128)
    stats = torch.full((tiles * 64,), float("nan"), device=DEV, dtype=torch.float32)
    for rep in range(2):                                      # second round: counter re-armed, same answer
        counter = torch.zeros(1, device=DEV, dtype=torch.int32)
        resid = resid0.clone()
        R = ops.decode_tile_rows(B)
        normed_t = torch.full((N // 64 * R * 64,), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.decode_gemm_resnorm(ops.tile_rows(x, R).view(-1), ops.tile_weight(w), resid, ln_w, normed_t, stats, counter, eps, K)
        torch.cuda.synchronize()
        normed = ops.untile_rows(normed_t.view(N // 64, R * 64), B, N)
        assert int(counter.item()) == tiles * 8
        # PyTorch restatement with HF's rounding points ([Q]:243,302-308, 258-263)
        y = _bf(x.float() @ w.float().t())
        xn = _bf(y.float() + resid0.float())
        var = xn.float().pow(2).mean(-1, keepdim=True)
        ref_normed = _bf(_bf(xn.float() * torch.rsqrt(var + eps)).float() * ln_w.float())
        assert float((resid == xn).float().mean()) > 0.97
        assert float((resid.float() - xn.float()).abs().max() / xn.float().abs().max()) < 8e-3
        assert not torch.isnan(normed.float()).any()
        assert float((normed.float() - ref_normed.float()).abs().max() / ref_normed.float().abs().max()) < 1.2e-2
        # per-op kernels with the same K partition
        nkb = -(-K // 64)
        per = -(-nkb // 8)
        splits = -(-nkb // per)
        part = ops.gemm_skinny(x, w, splits)
        resid_old, normed_old = resid0.clone(), torch.empty_like(resid0)
        ops.decode_residual_rmsnorm(part, splits, resid_old, ln_w, normed_old, eps)
        assert torch.equal(resid, resid_old)
        d = _ulp_diff(normed, normed_old)
        assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.995


@pytest.mark.parametrize("B,nq,nkv,ctx,splits", [(64, 12, 2, 1881, 2), (8, 12, 2, 700, 4), (3, 6, 1, 130, 2), (5, 12, 2, 64, 3), (2, 12, 2, 5000, 4),
                                                  (4, 12, 2, 3000, 8), (64, 12, 2, 1881, 1)])
def test_attention_from_bf16_qkv_and_cluster_merge(B, nq, nkv, ctx, splits, gen):
    """dots_attn_decode_qkv (bf16 q|k|v row, cluster merge, row-major or tiled output) == dots_attn_decode_fused (fp32 partials, combine kernel)."""
    ops = _ops()
    N = (nq + 2 * nkv) * 128
    ctx_max = (ctx + 64 + 63) // 64 * 64
    kc = _rand((B, nkv, ctx_max, 128), gen)                # random bits are as good in the tiled cache layout as in any other
    vc = _rand((B, nkv, ctx_max, 128), gen)
    lens = torch.randint(max(1, ctx - 90), ctx + 1, (B,), generator=gen, device=DEV)
    lens[0] = ctx
    pos = (lens - 1).to(torch.int32)
    ctx_len = lens.to(torch.int32)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(DEV)
    bias = _rand((N,), gen)
    part = torch.randn((3, B, N), generator=gen, device=DEV)
    qkv = _bf((part[0] + part[1]) + part[2] + bias.float())
    scale = 128 ** -0.5
    outs = []
    R = ops.decode_tile_rows(B)
    for mode in ("old_combine", "old_cluster", "new_cluster", "new_combine", "new_tiled_out"):
        k1, v1 = kc.clone(), vc.clone()
        out = torch.full((B, nq * 128), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.set_decode_cluster(mode.endswith("cluster"))
        try:
            po = torch.empty((B, nq, splits, 128), device=DEV, dtype=torch.float32)
            pml = torch.empty((B, nq, splits, 2), device=DEV, dtype=torch.float32)
            if mode.startswith("old"):
                ops.attn_decode_fused(part, 3, bias, pos, inv_freq, k1, v1, ctx_len, out, nq, nkv, ctx_max, splits, scale, po, pml)
            elif mode == "new_tiled_out":
                out_t = torch.full((nq * 2 * R * 64,), float("nan"), device=DEV, dtype=torch.bfloat16)
                ops.attn_decode_qkv(qkv, pos, inv_freq, k1, v1, ctx_len, out_t, nq, nkv, ctx_max, splits, scale, po, pml, out_tile_rows=R)
                out = ops.untile_rows(out_t.view(nq * 2, R * 64), B, nq * 128)
            else:
                ops.attn_decode_qkv(qkv, pos, inv_freq, k1, v1, ctx_len, out, nq, nkv, ctx_max, splits, scale, po, pml)
        finally:
            ops.set_decode_cluster(True)
        torch.cuda.synchronize()
        assert not torch.isnan(out.float()).any(), mode
        outs.append((out, k1, v1))
    for o, k1, v1 in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(k1, outs[0][1]) and torch.equal(v1, outs[0][2])
    # and against fp32 softmax attention over the appended cache
    out, k1, v1 = outs[2]
    for b in range(min(B, 4)):
        L = int(lens[b])
        q = _hf_rope_bf16(qkv[b:b + 1, : nq * 128].reshape(1, nq, 128), pos[b:b + 1], inv_freq)[0].float()      # [nq, 128]
        kk = ops.kv_untile(k1[b])[:, :L].float().repeat_interleave(nq // nkv, 0)
        vv = ops.kv_untile(v1[b])[:, :L].float().repeat_interleave(nq // nkv, 0)
        p = torch.softmax(torch.einsum("hd,hld->hl", q, kk) * scale, -1)
        ref = torch.einsum("hl,hld->hd", p, vv).reshape(-1)
        assert float((out[b].float() - ref).abs().max()) < 2e-2


def test_fault_injection_changes_the_attention_output(gen):
    """The test-only fault switch really perturbs the kernel (the parity tests rely on it to prove they can fail)."""
    ops = _ops()
    B, nq, nkv, ctx = 4, 12, 2, 300
    N = (nq + 2 * nkv) * 128
    ctx_max = 384
    kc, vc = _rand((B, nkv, ctx_max, 128), gen), _rand((B, nkv, ctx_max, 128), gen)
    pos = torch.full((B,), ctx - 1, device=DEV, dtype=torch.int32)
    ctx_len = pos + 1
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(DEV)
    qkv = _rand((B, N), gen)
    outs = []
    try:
        for code in (0, 1, 0):
            ops.debug_set_fault(code)
            out = torch.empty((B, nq * 128), device=DEV, dtype=torch.bfloat16)
            ops.attn_decode_qkv(qkv, pos, inv_freq, kc.clone(), vc.clone(), ctx_len, out, nq, nkv, ctx_max, 2, 128 ** -0.5)
            outs.append(out)
    finally:
        ops.debug_set_fault(0)
    assert torch.equal(outs[0], outs[2])
    assert float((outs[0].float() - outs[1].float()).abs().max()) > 5e-2


def test_engine_decode_modes_agree():
    """Whole engine, tiny config: the decode-layer variants ("tiled": 7 kernels over bulk-copied tiled operands, "fused": 5 kernels with
    cluster GEMMs, "perop": 7 kernels over tensor-map copies) give the same greedy ids on the `peaked` checkpoint and teacher-forced
    logits that agree to bf16 noise on `random` weights; graph replay == eager stepping in every mode."""
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    cfg = config.tiny()
    g = torch.Generator().manual_seed(5)
    grids = [(1, 8, 8), (1, 4, 12), (1, 8, 12)]
    pvs, rows = [], []
    for (_, h, w) in grids:
        pvs.append(torch.randn(h * w, cfg.vision.patch_dim, generator=g))
        rows.append(torch.cat([torch.randint(0, 2000, (6,), generator=g), torch.full((h * w // 4,), cfg.image_token_id),
                               torch.randint(0, 2000, (9,), generator=g)]))
    T = max(r.numel() for r in rows)
    ids = torch.zeros((3, T), dtype=torch.long)
    mask = torch.zeros((3, T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    pv, grid = torch.cat(pvs).to(DEV), torch.tensor(grids)
    for flavour in ("peaked", "random"):
        eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, flavour), DEV)
        res = {}
        for mode in ("perop", "tiled", "fused"):
            eng.decode_mode = mode
            assert eng._decode_plan(3)["mode"] == mode
            a = eng.generate(ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=40)
            b = eng.generate(ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=40, use_graph=False,
                             return_logits=True)
            assert torch.equal(a.sequences, b.sequences), mode
            res[mode] = (a.sequences.cpu(), b.logits.float().cpu())
        if flavour == "peaked":
            assert torch.equal(res["tiled"][0], res["perop"][0]) and torch.equal(res["fused"][0], res["perop"][0])
        # same partial sums in the same order, operands only fetched differently / finalize only launched differently: bit-identical
        assert torch.equal(res["tiled"][1], res["perop"][1]) and torch.equal(res["tiled"][0], res["perop"][0])
        forced = res["perop"][0][:, T:]
        eng.decode_mode = "fused"
        lf = eng.generate(ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=40, forced_ids=forced,
                          return_logits=True).logits.float().cpu()
        sd = float(res["perop"][1].std())
        assert float((lf - res["perop"][1]).abs().max()) / sd < 6e-2
