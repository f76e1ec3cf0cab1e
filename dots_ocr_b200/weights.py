"""Checkpoint tensors for the dots.ocr hot path.

No real ``weights/DotsOCR`` directory exists offline, so
the default checkpoint is synthetic and seeded.  Tensor names are the HF
checkpoint names the vLLM mapper expects
(``vllm/model_executor/models/dots_ocr.py:620-643``), so
``load_safetensors_dir`` can read a real directory with the same code path.

Two synthetic flavours:

* ``"random"``  -- N(0, 0.02) linears (``initializer_range``), norms near 1.  Greedy
  argmax over 152 k near-Gaussian logits has tiny top-1 margins, so this flavour
  is used for teacher-forced logit parity, not for free-running id equality.
* ``"peaked"``  -- same body, but ``embed_tokens`` has unit scale and
  ``lm_head[perm[t]] = embed_tokens[t]`` so that the residual stream's embedding
  component selects the next token with a margin far above bf16 noise.  Used
  for the 512-step bit-exact greedy-id checks (SURVEY.md §7.3 mitigation (b)).
"""
from __future__ import annotations

import os
from typing import Dict, Iterator, Tuple

import torch

from .config import DotsConfig


def _specs(cfg: DotsConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) for every checkpoint tensor, in a fixed order."""
    v, t = cfg.vision, cfg.text
    D = v.embed_dim
    yield "vision_tower.patch_embed.patchifier.proj.weight", (D, v.num_channels, v.patch_size, v.patch_size), "linear"
    yield "vision_tower.patch_embed.patchifier.proj.bias", (D,), "bias"
    yield "vision_tower.patch_embed.patchifier.norm.weight", (D,), "norm"
    for i in range(v.num_hidden_layers):
        p = f"vision_tower.blocks.{i}."
        yield p + "norm1.weight", (D,), "norm"
        yield p + "attn.qkv.weight", (3 * D, D), "linear"
        yield p + "attn.proj.weight", (D, D), "linear"
        yield p + "norm2.weight", (D,), "norm"
        yield p + "mlp.fc1.weight", (v.intermediate_size, D), "linear"
        yield p + "mlp.fc3.weight", (v.intermediate_size, D), "linear"
        yield p + "mlp.fc2.weight", (D, v.intermediate_size), "linear"
    yield "vision_tower.post_trunk_norm.weight", (D,), "norm"
    yield "vision_tower.merger.ln_q.weight", (D,), "norm"
    yield "vision_tower.merger.ln_q.bias", (D,), "bias"
    yield "vision_tower.merger.mlp.0.weight", (v.merge_dim, v.merge_dim), "linear"
    yield "vision_tower.merger.mlp.0.bias", (v.merge_dim,), "bias"
    yield "vision_tower.merger.mlp.2.weight", (v.hidden_size, v.merge_dim), "linear"
    yield "vision_tower.merger.mlp.2.bias", (v.hidden_size,), "bias"

    H = t.hidden_size
    yield "model.embed_tokens.weight", (t.vocab_size, H), "embed"
    for i in range(t.num_hidden_layers):
        p = f"model.layers.{i}."
        yield p + "input_layernorm.weight", (H,), "norm"
        yield p + "self_attn.q_proj.weight", (t.q_dim, H), "linear"
        yield p + "self_attn.q_proj.bias", (t.q_dim,), "bias"
        yield p + "self_attn.k_proj.weight", (t.kv_dim, H), "linear"
        yield p + "self_attn.k_proj.bias", (t.kv_dim,), "bias"
        yield p + "self_attn.v_proj.weight", (t.kv_dim, H), "linear"
        yield p + "self_attn.v_proj.bias", (t.kv_dim,), "bias"
        yield p + "self_attn.o_proj.weight", (H, t.q_dim), "linear"
        yield p + "post_attention_layernorm.weight", (H,), "norm"
        yield p + "mlp.gate_proj.weight", (t.intermediate_size, H), "linear"
        yield p + "mlp.up_proj.weight", (t.intermediate_size, H), "linear"
        yield p + "mlp.down_proj.weight", (H, t.intermediate_size), "linear"
    yield "model.norm.weight", (H,), "norm"
    yield "lm_head.weight", (t.vocab_size, H), "head"


def tensor_names(cfg: DotsConfig):
    return [n for n, _, _ in _specs(cfg)]


def param_count(cfg: DotsConfig) -> Dict[str, int]:
    vis = txt = 0
    for name, shape, _ in _specs(cfg):
        n = 1
        for s in shape:
            n *= s
        if name.startswith("vision_tower."):
            vis += n
        else:
            txt += n
    return {"vision": vis, "text": txt, "total": vis + txt}


def make_synthetic_checkpoint(cfg: DotsConfig, seed: int = 0, flavour: str = "random",
                              device: str | torch.device = "cpu",
                              dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint.  Values are drawn in fp32 on ``device`` and
    rounded once to ``dtype``; the fp32 oracle upcasts these same rounded values,
    so both sides see identical parameters."""
    assert flavour in ("random", "peaked")
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    std = 0.02
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in _specs(cfg):
        if kind == "norm":
            w = 1.0 + 0.1 * torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        elif kind == "bias":
            w = std * torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        elif kind == "embed" and flavour == "peaked":
            w = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        elif kind == "head" and flavour == "peaked":
            # next(t) = perm[t]; perm is a fixed odd-multiplier affine map (a bijection
            # on [0, V) when gcd(a, V) == 1), so no V-sized randperm is needed.
            V = shape[0]
            a = _coprime_multiplier(V)
            t = torch.arange(V, device=device, dtype=torch.int64)
            perm = (a * t + 12345) % V
            emb = out["model.embed_tokens.weight"].float()
            w = torch.empty(shape, device=device, dtype=torch.float32)
            w[perm] = emb * 0.05
            # burn the generator the same amount as the random flavour would not matter:
            # the head is the last tensor.
        else:
            w = std * torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        out[name] = w.to(dtype)
    return out


def _coprime_multiplier(V: int) -> int:
    import math
    a = 48271 % V
    if a < 2:
        a = 3
    while math.gcd(a, V) != 1:
        a += 1
    return a


def peaked_next_token(cfg: DotsConfig, token: int) -> int:
    """The successor map baked into the ``peaked`` flavour's lm_head."""
    V = cfg.text.vocab_size
    return (_coprime_multiplier(V) * token + 12345) % V


def load_safetensors_dir(path: str, device: str | torch.device = "cpu",
                         dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Read a real HF checkpoint directory (``./weights/DotsOCR`` in the reference,
    ``dots_ocr/parser.py:67``).  Alternate spellings the vLLM mapper accepts are
    normalised (``.attn.qkv_proj.`` -> ``.attn.qkv.``, ``.attn.out_proj.`` -> ``.attn.proj.``)."""
    from safetensors import safe_open
    out: Dict[str, torch.Tensor] = {}
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors files in {path}")
    for f in files:
        with safe_open(os.path.join(path, f), framework="pt", device=str(device)) as sf:
            for k in sf.keys():
                name = k.replace(".attn.qkv_proj.", ".attn.qkv.").replace(".attn.out_proj.", ".attn.proj.")
                out[name] = sf.get_tensor(k).to(dtype)
    return out


def validate_checkpoint(cfg: DotsConfig, ckpt: Dict[str, torch.Tensor], allow_extra: bool = True) -> None:
    """Every tensor the engine will read exists with the shape ``cfg`` implies; raise ``ValueError`` naming the
    first few offenders otherwise (a silent mis-shape would surface as a wrong pitch inside a kernel)."""
    missing, wrong = [], []
    expected = set()
    for name, shape, _ in _specs(cfg):
        expected.add(name)
        w = ckpt.get(name)
        if w is None:
            missing.append(name)
        elif tuple(w.shape) != tuple(shape):
            wrong.append(f"{name}: {tuple(w.shape)} != {tuple(shape)}")
    extra = [] if allow_extra else sorted(k for k in ckpt if k not in expected)
    if missing or wrong or extra:
        def head(xs):
            return ", ".join(xs[:6]) + (f", ... (+{len(xs) - 6})" if len(xs) > 6 else "")
        parts = []
        if missing:
            parts.append(f"{len(missing)} missing [{head(missing)}]")
        if wrong:
            parts.append(f"{len(wrong)} mis-shaped [{head(wrong)}]")
        if extra:
            parts.append(f"{len(extra)} unexpected [{head(extra)}]")
        raise ValueError("checkpoint does not match the configuration: " + "; ".join(parts))


def save_safetensors_dir(ckpt: Dict[str, torch.Tensor], path: str, shards: int = 1) -> None:
    """Write ``ckpt`` as ``model-0000i-of-0000n.safetensors`` files (test / tooling helper: the inverse of
    ``load_safetensors_dir``)."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    names = list(ckpt)
    per = (len(names) + shards - 1) // shards
    for i in range(shards):
        part = {k: ckpt[k].contiguous() for k in names[i * per:(i + 1) * per]}
        if part:
            save_file(part, os.path.join(path, f"model-{i + 1:05d}-of-{shards:05d}.safetensors"))
