"""tcgen05 attention kernel (dots_attn_varlen_fwd_tc) vs an fp32 PyTorch reference and vs the mma.sync kernel.
Reference semantics: flash_attn_varlen_func as called at [V] dots_ocr.py:304-310 (bidirectional, per image) and
HF sdpa causal GQA at [Q] modeling_qwen2.py:161-183."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_attn(q, k, v, causal, group):
    qf, kf, vf = q.float().transpose(0, 1), k.float().transpose(0, 1), v.float().transpose(0, 1)
    kf, vf = kf.repeat_interleave(group, 0), vf.repeat_interleave(group, 0)
    s = qf @ kf.transpose(1, 2) / math.sqrt(128)
    if causal:
        L = q.shape[0]
        s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=q.device), 1), float("-inf"))
    return (torch.softmax(s, -1) @ vf).transpose(0, 1)


@pytest.mark.parametrize("lens,hq,hkv,causal,scale_q", [
    ([64], 2, 2, False, 1.0), ([128], 1, 1, False, 1.0), ([256], 1, 1, False, 1.0), ([300], 2, 2, False, 1.0),
    ([100, 37, 256], 2, 2, False, 1.0), ([1369], 12, 12, False, 1.0), ([5476], 2, 2, False, 1.0),
    ([5476], 1, 1, False, 6.0),            # peaked scores: exercises the lazy-rescale path
    ([70, 1, 300], 6, 1, True, 1.0), ([1625], 12, 2, True, 1.0), ([1625, 900], 6, 1, True, 5.0)])
@pytest.mark.parametrize("impl", ["tc"])
def test_attn_tc(lens, hq, hkv, causal, scale_q, impl):
    from dots_ocr_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(1)
    T = sum(lens)
    group = hq // hkv
    qkv = torch.randn((T, (hq + 2 * hkv) * 128), generator=g, device=DEV)
    qkv[:, : hq * 128] *= scale_q
    qkv = qkv.to(torch.bfloat16)
    q, k, v = qkv[:, : hq * 128], qkv[:, hq * 128:(hq + hkv) * 128], qkv[:, (hq + hkv) * 128:]
    out = torch.full((T, hq * 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    ops.attn_varlen(q, k, v, out, cu, max(lens), hq, hkv, causal, 128 ** -0.5, impl=impl)
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any(), "rows left unwritten / NaN"
    a = 0
    for L in lens:
        ref = _ref_attn(q[a:a + L].reshape(L, hq, 128), k[a:a + L].reshape(L, hkv, 128), v[a:a + L].reshape(L, hkv, 128),
                        causal, group)
        got = out[a:a + L].reshape(L, hq, 128).float()
        err = float((got - ref).abs().max())
        assert err < 2e-2, (L, err)          # |v| ~ 1; P rounded to bf16 inside the kernel
        a += L
