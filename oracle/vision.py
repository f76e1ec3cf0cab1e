"""ORACLE (test infrastructure): plain-PyTorch DotsVisionTransformer forward.

Follows ``vllm/model_executor/models/dots_ocr.py`` (cited as [V] below) with the
rounding points of the HF eager path: every Linear rounds to the working dtype, RMSNorm
normalises in fp32 then casts then multiplies by the weight, 2-D RoPE is applied in fp32,
SwiGLU is ``silu(fc1 x) * fc3 x`` with both factors in the working dtype.
Runs in fp32 on CPU (ground truth) or bf16 on CUDA ("HF path on the same box").
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    # [V]:450,456,518 (RMSNorm); HF-style: fp32 normalise -> cast -> * weight
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return xf.to(x.dtype) * w


def vision_pos_ids(grid_thw: Sequence[Sequence[int]], merge: int) -> torch.Tensor:
    """(h, w) index of every token in processor order (2x2 merge blocks contiguous). [V]:536-560"""
    out = []
    for t, h, w in grid_thw:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w)
        hp = hp.reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1)
        wp = wp.reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        out.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
    return torch.cat(out, dim=0)


def rot_pos_emb(grid_thw, merge: int, head_dim: int, theta: float, device) -> torch.Tensor:
    """[S, head_dim/2] fp32 angles = [h*f_0..h*f_{n-1}, w*f_0..w*f_{n-1}]. [V]:163-174, 562-568"""
    dim = head_dim // 2
    inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
    max_grid = max(max(h, w) for _, h, w in grid_thw)
    seq = torch.arange(max_grid, dtype=torch.float)
    freqs = torch.outer(seq, inv_freq)                      # [max_grid, dim/2]
    pos = vision_pos_ids(grid_thw, merge)
    return freqs[pos].flatten(1).to(device)                 # [S, dim]


def apply_rope_fp32(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """NeoX rotate-half in fp32, x: [S, H, hd]; cos/sin: [S, hd/2].
    rotary_embedding/common.py:144-183 with enable_fp32_compute=True ([V]:272-275)."""
    dt = x.dtype
    xf = x.float()
    c = cos.float().unsqueeze(-2)
    s = sin.float().unsqueeze(-2)
    x1, x2 = torch.chunk(xf, 2, dim=-1)
    o1 = x1 * c - x2 * s
    o2 = x2 * c + x1 * s
    return torch.cat((o1, o2), dim=-1).to(dt)


class VisionOracle:
    def __init__(self, vcfg, ckpt: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu"):
        self.cfg = vcfg
        self.dtype = dtype
        self.device = torch.device(device)
        self.w = {k[len("vision_tower."):]: v.to(device=self.device, dtype=dtype)
                  for k, v in ckpt.items() if k.startswith("vision_tower.")}

    # -- pieces (exposed for op-level parity tests) ---------------------------------
    def patch_embed(self, pixel_values: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        x = pixel_values.to(self.dtype)                                   # [V]:586
        x = x.view(-1, c.num_channels, c.temporal_patch_size, c.patch_size, c.patch_size)[:, :, 0]
        x = F.conv2d(x, self.w["patch_embed.patchifier.proj.weight"],
                     self.w["patch_embed.patchifier.proj.bias"], stride=c.patch_size)
        x = x.view(-1, c.embed_dim)                                       # [V]:405-415
        return rms_norm(x, self.w["patch_embed.patchifier.norm.weight"], c.rms_norm_eps)

    def attention(self, i: int, x: torch.Tensor, cu: List[int], cos, sin) -> torch.Tensor:
        c = self.cfg
        S = x.shape[0]
        H, hd = c.num_attention_heads, c.head_dim
        qkv = F.linear(x, self.w[f"blocks.{i}.attn.qkv.weight"])         # [V]:287
        q, k, v = qkv.chunk(3, dim=-1)                                    # qwen2_vl.py:335
        q = apply_rope_fp32(q.reshape(S, H, hd), cos, sin)
        k = apply_rope_fp32(k.reshape(S, H, hd), cos, sin)
        v = v.reshape(S, H, hd)
        out = torch.empty_like(q)
        for a, b in zip(cu[:-1], cu[1:]):                                 # one bidirectional segment per image
            o = F.scaled_dot_product_attention(q[a:b].transpose(0, 1).unsqueeze(0),
                                               k[a:b].transpose(0, 1).unsqueeze(0),
                                               v[a:b].transpose(0, 1).unsqueeze(0))
            out[a:b] = o.squeeze(0).transpose(0, 1)
        return F.linear(out.reshape(S, H * hd), self.w[f"blocks.{i}.attn.proj.weight"])

    def mlp(self, i: int, x: torch.Tensor) -> torch.Tensor:
        p = f"blocks.{i}.mlp."
        h = F.silu(F.linear(x, self.w[p + "fc1.weight"])) * F.linear(x, self.w[p + "fc3.weight"])
        return F.linear(h, self.w[p + "fc2.weight"])                      # [V]:334-356

    def block(self, i: int, x, cu, cos, sin):
        c = self.cfg
        x = x + self.attention(i, rms_norm(x, self.w[f"blocks.{i}.norm1.weight"], c.rms_norm_eps), cu, cos, sin)
        x = x + self.mlp(i, rms_norm(x, self.w[f"blocks.{i}.norm2.weight"], c.rms_norm_eps))
        return x                                                           # [V]:458-473

    def merger(self, x: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        x = F.layer_norm(x, (c.embed_dim,), self.w["merger.ln_q.weight"], self.w["merger.ln_q.bias"],
                         c.merger_ln_eps)
        x = x.view(-1, c.merge_dim)
        x = F.linear(x, self.w["merger.mlp.0.weight"], self.w["merger.mlp.0.bias"])
        x = F.gelu(x)
        return F.linear(x, self.w["merger.mlp.2.weight"], self.w["merger.mlp.2.bias"])   # [V]:215-220

    # -- whole tower -------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, grid_thw, return_layers: bool = False):
        c = self.cfg
        grid = [list(map(int, g)) for g in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)]
        ang = rot_pos_emb(grid, c.spatial_merge_size, c.head_dim, c.rope_theta, self.device)
        cos, sin = ang.cos(), ang.sin()
        x = self.patch_embed(pixel_values.to(self.device))
        cu = [0]
        for t, h, w in grid:
            for _ in range(t):
                cu.append(cu[-1] + h * w)                                  # [V]:590-596
        layers = [x]
        for i in range(c.num_hidden_layers):
            x = self.block(i, x, cu, cos, sin)
            if return_layers:
                layers.append(x)
        x = rms_norm(x, self.w["post_trunk_norm.weight"], c.rms_norm_eps)  # [V]:607-608
        x = self.merger(x)
        return (x, layers) if return_layers else x
