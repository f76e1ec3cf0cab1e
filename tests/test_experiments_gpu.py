"""Kernels kept as documented negative results (csrc/experiments/) still compute the right thing.
Skipped unless the library was built with DOTS_BUILD_EXPERIMENTS=1 (they are not part of the product library)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from dots_ocr_b200 import ops
    if not ops.has_experiments():
        pytest.skip("experiment kernels are not in this build (DOTS_BUILD_EXPERIMENTS=1)")
    return ops


def _bf(x):
    return x.to(torch.bfloat16)


def _rand(shape, gen, scale=1.0):
    return _bf(torch.randn(shape, generator=gen, device=DEV) * scale)


@pytest.fixture(scope="module")
def gen():
    g = torch.Generator(device=DEV)
    g.manual_seed(1234)
    return g


@pytest.mark.parametrize("B,H,I,QKV,so,sd,sq", [(64, 1536, 8960, 2048, 12, 12, 8), (7, 768, 1024, 1024, 12, 16, 12), (33, 1536, 4224, 2048, 6, 11, 4),
                                                (1, 1536, 8960, 2048, 12, 12, 8)])
def test_decode_chain_matches_per_op_kernels(B, H, I, QKV, so, sd, sq, gen):
    """The persistent per-layer decode kernel is bit-identical to the sequence of per-op kernels it replaces."""
    from dots_ocr_b200.engine import _interleave_gate_up
    ops = _ops()
    eps = 1e-6
    attn = _rand((B, H), gen)
    w_o, w_down, w_qkv = _rand((H, H), gen, 0.03), _rand((H, I), gen, 0.02), _rand((QKV, H), gen, 0.03)
    w_gu = _interleave_gate_up(_rand((I, H), gen, 0.03), _rand((I, H), gen, 0.03))
    ln_mid, ln_next = _bf(1 + 0.1 * torch.randn(H, generator=gen, device=DEV)), _bf(1 + 0.1 * torch.randn(H, generator=gen, device=DEV))
    resid0 = _rand((B, H), gen)
    nmax = max(so * H, sd * H, sq * QKV) * B
    for with_qkv in (True, False):
        # reference: per-op kernels
        part = torch.zeros(nmax, device=DEV, dtype=torch.float32)
        resid, normed, act = resid0.clone(), torch.empty_like(resid0), torch.empty((B, I), device=DEV, dtype=torch.bfloat16)
        ops.gemm_skinny(attn, w_o, so, partial=part)
        ops.decode_residual_rmsnorm(part, so, resid, ln_mid, normed, eps)
        ops.gemm_skinny_swiglu(normed, w_gu, act)
        ops.gemm_skinny(act, w_down, sd, partial=part)
        ops.decode_residual_rmsnorm(part, sd, resid, ln_next, normed, eps)
        if with_qkv:
            ops.gemm_skinny(normed, w_qkv, sq, partial=part)
        # chain kernel
        part2 = torch.zeros(nmax, device=DEV, dtype=torch.float32)
        resid2, normed2, act2 = resid0.clone(), torch.full_like(resid0, float("nan")), torch.full((B, I), float("nan"), device=DEV, dtype=torch.bfloat16)
        counters = torch.zeros(8, device=DEV, dtype=torch.int32)
        ops.decode_chain(attn, w_o, w_gu, w_down, w_qkv if with_qkv else None, part2, resid2, normed2, act2, ln_mid, ln_next, counters,
                         so, sd, sq, eps)
        torch.cuda.synchronize()
        assert torch.equal(act, act2)
        assert torch.equal(resid, resid2)
        assert torch.equal(normed, normed2)
        if with_qkv:
            assert torch.equal(part[: sq * B * QKV], part2[: sq * B * QKV])


@pytest.mark.parametrize("lens,hq,hkv,causal", [([300], 2, 2, False), ([100, 37, 256], 2, 2, False), ([5476], 2, 2, False),
                                                ([1625, 900], 6, 1, True)])
def test_attn_pair_matches_single_cta_kernel(lens, hq, hkv, causal):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(1)
    T = sum(lens)
    qkv = torch.randn((T, (hq + 2 * hkv) * 128), generator=g, device=DEV).to(torch.bfloat16)
    q, k, v = qkv[:, : hq * 128], qkv[:, hq * 128:(hq + hkv) * 128], qkv[:, (hq + hkv) * 128:]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    a = torch.full((T, hq * 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    b = torch.full_like(a, float("nan"))
    ops.attn_varlen(q, k, v, a, cu, max(lens), hq, hkv, causal, 128 ** -0.5, impl="tc")
    ops.attn_varlen(q, k, v, b, cu, max(lens), hq, hkv, causal, 128 ** -0.5, impl="pair")
    assert float((a.float() - b.float()).abs().max()) < 2e-2
