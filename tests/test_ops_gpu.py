"""Op-level parity: every C-ABI kernel against a plain PyTorch restatement of the same op with the
HF rounding points (oracle-side code lives in oracle/ and in these reference lambdas).
Tolerances: integer/index work is bit-exact; bf16 outputs may differ from the torch reference by
one bf16 ulp on a small fraction of elements (different fp32 accumulation order) -- the checks state
max relative error and, where the op is a pure elementwise/rounding op, demand exact equality."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from dots_ocr_b200 import ops
    return ops


def _rel_err(out, ref):
    out, ref = out.float(), ref.float()
    return float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-6))


def _frac_exact(out, ref):
    return float((out == ref).float().mean())


def _bf(x):
    return x.to(torch.bfloat16)


def _rand(shape, gen, scale=1.0):
    return _bf(torch.randn(shape, generator=gen, device=DEV) * scale)


@pytest.fixture(scope="module")
def gen():
    g = torch.Generator(device=DEV)
    g.manual_seed(1234)
    return g


GEMM_SHAPES = [(128, 128, 64), (128, 256, 128), (300, 384, 192), (1000, 1536, 1536), (5476, 4608, 1536),
               (77, 136, 72), (1369, 768, 1024), (4096, 1536, 4224), (2048, 2048, 640)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store(M, N, K, gen):
    ops = _ops()
    a, w = _rand((M, K), gen), _rand((N, K), gen, 0.05)
    out = ops.gemm(a, w)
    ref = _bf(a.float() @ w.float().t())
    assert _rel_err(out, ref) < 8e-3, _rel_err(out, ref)
    assert _frac_exact(out, ref) > 0.98


@pytest.mark.parametrize("M,N,K", [(300, 384, 192), (5476, 1536, 1536), (1369, 6144, 6144)])
def test_gemm_bias_and_gelu(M, N, K, gen):
    ops = _ops()
    a, w, b = _rand((M, K), gen), _rand((N, K), gen, 0.03), _rand((N,), gen, 0.5)
    acc = a.float() @ w.float().t() + b.float()
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS, bias=b)
    assert _rel_err(out, _bf(acc)) < 8e-3
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=b)
    ref = torch.nn.functional.gelu(_bf(acc))
    assert _rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("M,N,K", [(300, 384, 192), (5476, 1536, 4224)])
def test_gemm_residual_inplace(M, N, K, gen):
    ops = _ops()
    a, w, r = _rand((M, K), gen), _rand((N, K), gen, 0.03), _rand((M, N), gen)
    ref = _bf(_bf(a.float() @ w.float().t()).float() + r.float())
    x = r.clone()
    ops.gemm(a, w, out=x, epilogue=ops.EPI_RESIDUAL, residual=x)
    assert _rel_err(x, ref) < 8e-3
    assert _frac_exact(x, ref) > 0.98


@pytest.mark.parametrize("M,I,K", [(300, 512, 256), (5476, 4224, 1536), (1625, 8960, 1536)])
def test_gemm_swiglu(M, I, K, gen):
    from dots_ocr_b200.engine import _interleave_gate_up
    ops = _ops()
    a, wg, wu = _rand((M, K), gen), _rand((I, K), gen, 0.03), _rand((I, K), gen, 0.03)
    w = _interleave_gate_up(wg, wu)
    out = ops.gemm(a, w, epilogue=ops.EPI_SWIGLU)
    g, u = _bf(a.float() @ wg.float().t()), _bf(a.float() @ wu.float().t())
    ref = torch.nn.functional.silu(g) * u
    assert out.shape == (M, I)
    assert _rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("B,I,K", [(1, 256, 512), (7, 1024, 1536), (64, 8960, 1536), (33, 4224, 1536), (130, 512, 768)])
def test_gemm_skinny_swiglu(B, I, K, gen):
    """Fused decode gate|up + SwiGLU == unfused skinny GEMM partials + decode_swiglu (same rounding points), and ~ fp32 reference."""
    from dots_ocr_b200.engine import _interleave_gate_up
    ops = _ops()
    x, wg, wu = _rand((B, K), gen), _rand((I, K), gen, 0.03), _rand((I, K), gen, 0.03)
    w = _interleave_gate_up(wg, wu)
    act = torch.full((B, I), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm_skinny_swiglu(x, w, act)
    part = ops.gemm_skinny(x, w, 1)
    act2 = torch.empty_like(act)
    ops.decode_swiglu(part, 1, act2)
    assert torch.equal(act, act2)
    g, u = _bf(x.float() @ wg.float().t()), _bf(x.float() @ wu.float().t())
    assert _rel_err(act, torch.nn.functional.silu(g) * u) < 1e-2


def test_attn_decode_fused_matches_unfused(gen):
    """QKV finalize fused into the decode attention kernel == decode_qkv_rope_append + attn_decode, bit for bit."""
    ops = _ops()
    nq, nkv = 12, 2
    N = (nq + 2 * nkv) * 128
    for (B, splits_qkv, n_splits, ctxs) in [(3, 8, 1, [5, 130, 64]), (2, 3, 4, [700, 2137]), (64, 8, 1, None), (1, 1, 16, [1])]:
        if ctxs is None:
            ctxs = [int(v) for v in torch.randint(1, 1900, (B,), generator=torch.Generator().manual_seed(5))]
        ctx_max = (max(ctxs) + 9 + 63) // 64 * 64
        part = torch.randn((splits_qkv, B, N), generator=gen, device=DEV)
        bias = _rand((N,), gen, 0.1)
        ctx = torch.tensor(ctxs, dtype=torch.int32, device=DEV)
        pos = ctx - 1
        inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(DEV)
        kc0, vc0 = _rand((B, nkv, ctx_max, 128), gen), _rand((B, nkv, ctx_max, 128), gen)
        # unfused
        kc, vc = kc0.clone(), vc0.clone()
        q = torch.empty((B, nq * 128), device=DEV, dtype=torch.bfloat16)
        ops.decode_qkv_rope_append(part, splits_qkv, bias, pos, inv_freq, q, kc, vc, ctx_max, nq, nkv)
        ref = torch.empty_like(q)
        ops.attn_decode(q, kc, vc, ctx, ref, nq, nkv, ctx_max, n_splits, 128 ** -0.5)
        # fused
        kc2, vc2 = kc0.clone(), vc0.clone()
        out = torch.full_like(q, float("nan"))
        ops.attn_decode_fused(part, splits_qkv, bias, pos, inv_freq, kc2, vc2, ctx, out, nq, nkv, ctx_max, n_splits, 128 ** -0.5)
        assert torch.equal(kc, kc2) and torch.equal(vc, vc2)
        assert torch.equal(out, ref), (B, float((out.float() - ref.float()).abs().max()))


@pytest.mark.parametrize("B,N,K,splits", [(1, 1024, 768, 12), (7, 2048, 1536, 8), (64, 1536, 8960, 12), (64, 17920, 1536, 1),
                                          (33, 1536, 1536, 12), (130, 2048, 1536, 4)])
def test_gemm_skinny_partials(B, N, K, splits, gen):
    ops = _ops()
    x, w = _rand((B, K), gen), _rand((N, K), gen, 0.05)
    part = ops.gemm_skinny(x, w, splits)
    assert part.shape == (splits, B, N)
    ref = x.float() @ w.float().t()
    got = part.sum(0)
    assert _rel_err(got, ref) < 1e-4
    # each split is its own K slice
    kb = -(-K // 64)
    per = -(-kb // splits) * 64
    for s in range(splits):
        sl = slice(s * per, min(K, (s + 1) * per))
        assert _rel_err(part[s], x[:, sl].float() @ w[:, sl].float().t()) < 1e-4


@pytest.mark.parametrize("B,N,K", [(1, 2048, 768), (64, 151936, 1536), (5, 2048, 1536)])
def test_gemm_skinny_bf16(B, N, K, gen):
    ops = _ops()
    x, w = _rand((B, K), gen), _rand((N, K), gen, 0.05)
    out = torch.empty((B, N), device=DEV, dtype=torch.bfloat16)
    ops.gemm_skinny(x, w, 1, out_bf16=out)
    ref = _bf(x.float() @ w.float().t())
    assert _rel_err(out, ref) < 8e-3
    assert _frac_exact(out, ref) > 0.98


@pytest.mark.parametrize("rows,cols", [(1, 768), (1000, 1536), (5476, 1536), (37, 256)])
def test_rmsnorm_exact(rows, cols, gen):
    ops = _ops()
    x, w = _rand((rows, cols), gen, 3.0), _bf(1 + 0.1 * torch.randn(cols, generator=gen, device=DEV))
    out = ops.rmsnorm(x, w, 1e-5)
    xf = x.float()
    ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16) * w
    assert _rel_err(out, ref) < 8e-3
    assert _frac_exact(out, ref) > 0.999


def test_layernorm(gen):
    ops = _ops()
    x = _rand((1369, 1536), gen, 2.0)
    w, b = _bf(1 + 0.1 * torch.randn(1536, generator=gen, device=DEV)), _rand((1536,), gen, 0.1)
    out = ops.layernorm(x, w, b, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (1536,), w, b, 1e-6)
    assert _rel_err(out, ref) < 8e-3
    assert _frac_exact(out, ref) > 0.99


def test_cast_pad(gen):
    ops = _ops()
    x = torch.randn((1000, 588), generator=gen, device=DEV)
    out = ops.cast_pad(x, 640)
    assert torch.equal(out[:, :588], x.to(torch.bfloat16))
    assert torch.count_nonzero(out[:, 588:]) == 0
    out2 = ops.cast_pad(x.to(torch.bfloat16), 640)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("grids", [[(1, 8, 8)], [(1, 74, 74)], [(1, 4, 6), (1, 10, 2), (2, 6, 6)]])
def test_vit_rope(grids, gen):
    from oracle.vision import rot_pos_emb, apply_rope_fp32
    ops = _ops()
    S = sum(t * h * w for t, h, w in grids)
    seqlens = [h * w for t, h, w in grids for _ in range(t)]
    ghw = [[h, w] for t, h, w in grids for _ in range(t)]
    cu = torch.tensor([0] + list(torch.tensor(seqlens).cumsum(0)), dtype=torch.int32, device=DEV)
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float) / 64))).to(DEV)
    cos, sin = ops.vit_rope_table(cu, torch.tensor(ghw, dtype=torch.int32, device=DEV), inv_freq, 2, S)
    ang = rot_pos_emb([list(g) for g in grids], 2, 128, 10000.0, DEV)
    assert torch.allclose(cos, ang.cos(), atol=2e-6, rtol=0) and torch.allclose(sin, ang.sin(), atol=2e-6, rtol=0)
    heads = 3
    qkv = _rand((S, 3 * heads * 128), gen)
    ref_q = apply_rope_fp32(qkv[:, : heads * 128].reshape(S, heads, 128), cos, sin).reshape(S, -1)
    ref_k = apply_rope_fp32(qkv[:, heads * 128: 2 * heads * 128].reshape(S, heads, 128), cos, sin).reshape(S, -1)
    ref_v = qkv[:, 2 * heads * 128:].clone()
    ops.vit_rope_apply(qkv, heads, cos, sin)
    assert torch.equal(qkv[:, : heads * 128], ref_q)
    assert torch.equal(qkv[:, heads * 128: 2 * heads * 128], ref_k)
    assert torch.equal(qkv[:, 2 * heads * 128:], ref_v)


def _hf_rope_bf16(x, pos, inv_freq):
    # modeling_qwen2.py:102-146: cos/sin in bf16, q*cos + rotate_half(q)*sin in bf16
    freqs = pos.float()[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(torch.bfloat16)[:, None, :], emb.sin().to(torch.bfloat16)[:, None, :]
    x1, x2 = x[..., :64], x[..., 64:]
    rot = torch.cat((-x2, x1), dim=-1)
    return (x * cos) + (rot * sin)


def test_llm_rope_kv_append(gen):
    ops = _ops()
    nq, nkv, B, ctx_max = 6, 1, 3, 64
    lens = [5, 17, 9]
    T = sum(lens)
    qkv = _rand((T, (nq + 2 * nkv) * 128), gen)
    pos = torch.cat([torch.arange(l) for l in lens]).to(torch.int32).to(DEV)
    seq = torch.cat([torch.full((l,), i) for i, l in enumerate(lens)]).to(torch.int32).to(DEV)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(DEV)
    kc = torch.zeros((B, nkv, ctx_max, 128), device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    ref_q = _hf_rope_bf16(qkv[:, : nq * 128].reshape(T, nq, 128), pos, inv_freq).reshape(T, -1)
    ref_k = _hf_rope_bf16(qkv[:, nq * 128:(nq + nkv) * 128].reshape(T, nkv, 128), pos, inv_freq)
    ref_v = qkv[:, (nq + nkv) * 128:].reshape(T, nkv, 128).clone()
    ops.llm_rope_kv_append(qkv, nq, nkv, pos, seq, inv_freq, kc, vc, ctx_max)
    kc, vc = ops.kv_untile(kc), ops.kv_untile(vc)          # the cache is stored in 64-key tiles (include/dots_ocr_b200.h)
    assert _frac_exact(qkv[:, : nq * 128], ref_q) > 0.999 and _rel_err(qkv[:, : nq * 128], ref_q) < 8e-3
    for t in range(T):
        b, p = int(seq[t]), int(pos[t])
        assert _rel_err(kc[b, :, p], ref_k[t]) < 8e-3
        assert torch.equal(vc[b, :, p], ref_v[t])


def _ref_attn(q, k, v, causal, group):
    # q [L, Hq, 128], k/v [L, Hkv, 128] -> fp32 reference with bf16-rounded inputs
    qf, kf, vf = q.float().transpose(0, 1), k.float().transpose(0, 1), v.float().transpose(0, 1)
    kf, vf = kf.repeat_interleave(group, 0), vf.repeat_interleave(group, 0)
    s = qf @ kf.transpose(1, 2) / math.sqrt(128)
    if causal:
        L = q.shape[0]
        s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=q.device), 1), float("-inf"))
    return (torch.softmax(s, -1) @ vf).transpose(0, 1)


@pytest.mark.parametrize("lens,hq,hkv,causal", [([64], 2, 2, False), ([100, 37, 256], 2, 2, False), ([1369], 12, 12, False),
                                                  ([5476], 2, 2, False), ([70, 1, 300], 6, 1, True), ([1625], 12, 2, True)])
def test_attn_varlen(lens, hq, hkv, causal, gen):
    ops = _ops()
    T = sum(lens)
    group = hq // hkv
    qkv = _rand((T, (hq + 2 * hkv) * 128), gen)
    q, k, v = qkv[:, : hq * 128], qkv[:, hq * 128:(hq + hkv) * 128], qkv[:, (hq + hkv) * 128:]
    out = torch.empty((T, hq * 128), device=DEV, dtype=torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    ops.attn_varlen(q, k, v, out, cu, max(lens), hq, hkv, causal, 128 ** -0.5)
    a = 0
    for L in lens:
        ref = _ref_attn(q[a:a + L].reshape(L, hq, 128), k[a:a + L].reshape(L, hkv, 128), v[a:a + L].reshape(L, hkv, 128),
                        causal, group)
        got = out[a:a + L].reshape(L, hq, 128).float()
        err = float((got - ref).abs().max())
        assert err < 2e-2, (L, err)          # |v| ~ 1; P rounded to bf16 inside the kernel
        a += L


@pytest.mark.parametrize("B,hq,hkv,ctxs,splits", [(1, 6, 1, [1], 1), (3, 12, 2, [5, 130, 64], 1), (3, 12, 2, [5, 130, 64], 4),
                                                   (64, 12, 2, None, 3), (2, 6, 1, [2137, 700], 16)])
def test_attn_decode(B, hq, hkv, ctxs, splits, gen):
    ops = _ops()
    if ctxs is None:
        ctxs = [int(x) for x in torch.randint(1, 2000, (B,), generator=torch.Generator().manual_seed(3))]
    ctx_max = (max(ctxs) + 7 + 63) // 64 * 64
    q = _rand((B, hq * 128), gen)
    kc, vc = _rand((B, hkv, ctx_max, 128), gen), _rand((B, hkv, ctx_max, 128), gen)
    ctx = torch.tensor(ctxs, dtype=torch.int32, device=DEV)
    out = torch.empty_like(q)
    ops.attn_decode(q, ops.kv_tile(kc), ops.kv_tile(vc), ctx, out, hq, hkv, ctx_max, splits, 128 ** -0.5)
    group = hq // hkv
    for b in range(B):
        L = ctxs[b]
        qq = q[b].reshape(hq, 128).float()
        kk = kc[b, :, :L].float().repeat_interleave(group, 0)
        vv = vc[b, :, :L].float().repeat_interleave(group, 0)
        s = torch.einsum("hd,hld->hl", qq, kk) / math.sqrt(128)
        ref = torch.einsum("hl,hld->hd", torch.softmax(s, -1), vv)
        err = float((out[b].reshape(hq, 128).float() - ref).abs().max())
        assert err < 2e-2, (b, L, err)


def test_embed_scatter_and_slots(gen):
    ops = _ops()
    V, H, T = 2048, 768, 5000
    table, img_tok = _rand((V, H), gen), 2040
    ids = torch.randint(0, 2000, (T,), device=DEV)
    ids[100:1500] = img_tok
    ids[3000:3007] = img_tok
    n_img = int((ids == img_tok).sum())
    img = _rand((n_img, H), gen)
    slots, count = ops.image_slots(ids, img_tok)
    assert int(count) == n_img
    ref_slots = torch.where(ids == img_tok, (ids == img_tok).cumsum(0) - 1, torch.full_like(ids, -1)).int()
    assert torch.equal(slots, ref_slots)
    out = ops.embed_scatter(ids, slots, table, img)
    emb = table[ids]
    mask = (ids == img_tok)
    ref = emb.masked_scatter(mask.unsqueeze(-1).expand_as(emb), img)
    assert torch.equal(out, ref)


def test_argmax_advance(gen):
    ops = _ops()
    B, V = 5, 151936
    logits = _rand((B, V), gen)
    logits[1, 777] = logits[1, 90000] = 50.0          # tie -> lowest index
    logits[2, 151935] = 60.0
    nxt = torch.zeros(B, device=DEV, dtype=torch.int64)
    out_ids = torch.full((B, 4), -7, device=DEV, dtype=torch.int64)
    step = torch.zeros(B, device=DEV, dtype=torch.int32)
    pos = torch.arange(B, device=DEV, dtype=torch.int32)
    ctx = pos + 1
    fin = torch.zeros(B, device=DEV, dtype=torch.int32)
    fin[4] = 1
    eos = int(logits[3].float().argmax())
    ops.argmax_advance(logits, nxt, out_ids, step, pos, ctx, fin, stop_ids=(eos, 151935), pad_id=11)      # two stop ids: rows 2 and 3 stop
    ref = logits.float().argmax(-1)
    assert nxt[0] == ref[0] and nxt[1] == 777 and nxt[2] == 151935 and nxt[3] == eos and nxt[4] == 11
    assert fin.tolist() == [0, 0, 1, 1, 1]
    assert torch.equal(out_ids[:, 0], nxt) and step.tolist() == [1] * B
    with pytest.raises(ValueError):
        ops.argmax_advance(logits, nxt, out_ids, step, pos, ctx, fin, stop_ids=tuple(range(ops.MAX_STOP_IDS + 1)), pad_id=11)
    assert pos.tolist() == [1, 2, 3, 4, 5] and ctx.tolist() == [2, 3, 4, 5, 6]


def test_decode_finalize_kernels(gen):
    from dots_ocr_b200.engine import _interleave_gate_up
    ops = _ops()
    B, H, I, S = 9, 1536, 8960, 5
    eps = 1e-6
    # residual + rmsnorm
    part = torch.randn((S, B, H), generator=gen, device=DEV)
    resid = _rand((B, H), gen)
    w = _bf(1 + 0.1 * torch.randn(H, generator=gen, device=DEV))
    r0 = resid.clone()
    normed = torch.empty_like(resid)
    ops.decode_residual_rmsnorm(part, S, resid, w, normed, eps)
    acc = part[0].clone()
    for s in range(1, S):
        acc = acc + part[s]
    x = _bf(_bf(acc).float() + r0.float())
    assert torch.equal(resid, x)
    xf = x.float()
    ref_n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(torch.bfloat16) * w
    assert _frac_exact(normed, ref_n) > 0.999
    # swiglu (interleaved layout)
    part = torch.randn((2, B, 2 * I), generator=gen, device=DEV)
    act = torch.empty((B, I), device=DEV, dtype=torch.bfloat16)
    ops.decode_swiglu(part, 2, act)
    tot = (part[0] + part[1]).view(B, I // 64, 2, 64)
    g, u = _bf(tot[:, :, 0].reshape(B, I)), _bf(tot[:, :, 1].reshape(B, I))
    ref = torch.nn.functional.silu(g) * u
    assert _frac_exact(act, ref) > 0.999 and _rel_err(act, ref) < 8e-3
    # embed + rmsnorm
    table = _rand((2048, H), gen)
    ids = torch.randint(0, 2048, (B,), device=DEV)
    ops.decode_embed_rmsnorm(ids, table, w, resid, normed, eps)
    assert torch.equal(resid, table[ids])
    xf = table[ids].float()
    ref_n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(torch.bfloat16) * w
    assert _frac_exact(normed, ref_n) > 0.999
    # qkv finalize + rope + append
    nq, nkv, ctx_max = 12, 2, 64
    N = (nq + 2 * nkv) * 128
    part = torch.randn((3, B, N), generator=gen, device=DEV)
    bias = _rand((N,), gen, 0.1)
    pos = torch.randint(0, ctx_max, (B,), device=DEV, dtype=torch.int32)
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).to(DEV)
    q_out = torch.empty((B, nq * 128), device=DEV, dtype=torch.bfloat16)
    kc = torch.zeros((B, nkv, ctx_max, 128), device=DEV, dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    ops.decode_qkv_rope_append(part, 3, bias, pos, inv_freq, q_out, kc, vc, ctx_max, nq, nkv)
    kc, vc = ops.kv_untile(kc), ops.kv_untile(vc)
    qkv = _bf((part[0] + part[1]) + part[2] + bias.float())
    ref_q = _hf_rope_bf16(qkv[:, : nq * 128].reshape(B, nq, 128), pos, inv_freq).reshape(B, -1)
    ref_k = _hf_rope_bf16(qkv[:, nq * 128:(nq + nkv) * 128].reshape(B, nkv, 128), pos, inv_freq)
    assert _frac_exact(q_out, ref_q) > 0.999
    for b in range(B):
        assert _frac_exact(kc[b, :, int(pos[b])], ref_k[b]) > 0.99
        assert torch.equal(vc[b, :, int(pos[b])], qkv[b, (nq + nkv) * 128:].reshape(nkv, 128))


@pytest.mark.parametrize("M,N,K,epi", [(20000, 1536, 1536, "store"), (43808, 4608, 1536, "store"), (43808, 1536, 1536, "res"),
                                        (26000, 2048, 1536, "bias"), (20000, 2048, 640, "gelu"), (43808, 8448, 1536, "swiglu"),
                                        (33111, 1536, 4224, "res")])
def test_gemm_cta_pair_matches_single_cta(M, N, K, epi, gen):
    """cta_group::2 kernel (256 x 256 tiles on CTA pairs) vs the one-CTA kernel on the same inputs: same MMA k-order, so equal."""
    ops = _ops()
    a, w = _rand((M, K), gen), _rand((N, K), gen, 0.03)
    bias = _rand((N,), gen) if epi in ("bias", "gelu") else None
    res0 = _rand((M, N), gen) if epi == "res" else None
    e = {"store": ops.EPI_STORE, "bias": ops.EPI_BIAS, "gelu": ops.EPI_BIAS_GELU, "res": ops.EPI_RESIDUAL, "swiglu": ops.EPI_SWIGLU}[epi]
    outs = []
    try:
        for pair in (False, True):
            ops.set_gemm_pair(pair)
            out = torch.full((M, N // 2 if epi == "swiglu" else N), float("nan"), device=DEV, dtype=torch.bfloat16)
            ops.gemm(a, w, out=out, epilogue=e, bias=bias, residual=res0)
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        ops.set_gemm_pair(True)         # library default
    assert not torch.isnan(outs[1].float()).any()
    assert torch.equal(outs[0], outs[1]), float((outs[0].float() - outs[1].float()).abs().max())


@pytest.mark.parametrize("H,W", [(56, 84), (1036, 1036), (280, 1008)])
def test_patchify_u8_matches_host_processor(H, W):
    """GPU rescale + normalise + patchify == host fp32 processor followed by cast_pad, bit for bit."""
    from dots_ocr_b200.processing import preprocess_image, CLIP_MEAN, CLIP_STD
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    pv, grid = preprocess_image(img.numpy(), min_pixels=H * W, max_pixels=H * W)       # no resize: sizes are multiples of 28
    assert grid.tolist() == [[1, H // 14, W // 14]]
    ref = ops.cast_pad(pv.to(DEV), 640)
    mean255 = (torch.tensor(CLIP_MEAN, dtype=torch.float32) * 255.0).tolist()
    std255 = (torch.tensor(CLIP_STD, dtype=torch.float32) * 255.0).tolist()
    got = ops.patchify_u8(img.to(DEV), 14, 2, mean255, std255, 640)
    assert got.shape == ref.shape
    assert torch.equal(got, ref)


@pytest.mark.parametrize("H,W,rh,rw", [(64, 80, 56, 84), (1024, 1024, 1036, 1036), (300, 200, 140, 84), (257, 311, 252, 308), (90, 64, 90, 56),
                                       (61, 70, 56, 70), (2250, 1700, 2240, 1708)])
def test_resize_bicubic_u8_equals_the_cpu_processor(H, W, rh, rw):
    """dots_resize_bicubic_u8 == torchvision's uint8 bicubic + antialias resize (the stock image processor's), bit for bit."""
    import torchvision.transforms.v2.functional as tvF
    from torchvision.transforms import InterpolationMode
    ops = _ops()
    g = torch.Generator().manual_seed(H + W)
    img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    ref = tvF.resize(img.permute(2, 0, 1).contiguous(), [rh, rw], interpolation=InterpolationMode.BICUBIC, antialias=True).permute(1, 2, 0)
    got = ops.resize_u8(img.to(DEV), rh, rw).cpu()
    assert torch.equal(got, ref.contiguous())


def test_gpu_image_processor_equals_host_processor():
    """Raw uint8 page of arbitrary size -> resize + rescale + normalise + patchify on the GPU == processing.preprocess_image on the host
    followed by the bf16 cast (the pixel_values the ViT's patch-embed GEMM consumes), bit for bit."""
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    from dots_ocr_b200.processing import preprocess_image
    ops = _ops()
    cfg = config.tiny()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "random"), DEV)
    g = torch.Generator().manual_seed(3)
    for (H, W) in [(100, 130), (224, 112), (75, 300)]:
        img = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
        pv, grid = preprocess_image(img.numpy())
        a = eng.encode_images(pv.to(DEV), grid)
        b = eng.encode_pages_u8([img.to(DEV)])
        assert torch.equal(a, b), (H, W)


@pytest.mark.parametrize("M,heads", [(43808, 12), (5476, 12), (300, 2), (20000, 12)])
def test_gemm_rope_epilogue_matches_gemm_then_rope(M, heads, gen):
    """ViT q|k|v projection with the 2-D rotary embedding in the GEMM epilogue == GEMM + dots_vit_rope_apply, bit for bit (large M runs
    the CTA-pair kernel with the fused epilogue, small M falls back to the two-kernel path inside the library)."""
    ops = _ops()
    D, K = heads * 128, 1536 if heads == 12 else 256
    a, w = _rand((M, K), gen), _rand((3 * D, K), gen, 0.03)
    ang = torch.rand((M, 64), generator=gen, device=DEV) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    ref = ops.gemm(a, w)
    ops.vit_rope_apply(ref, heads, cos, sin)
    out = torch.full((M, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm_rope(a, w, out, cos, sin, 2 * D)
    assert torch.equal(out, ref)
