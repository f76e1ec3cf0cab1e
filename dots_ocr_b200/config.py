"""Architecture constants for the dots.ocr page-parsing hot path.

The reference repo ships no model code; the numbers here follow
the in-container mirrors the survey cites:

* vision tower  -- ``vllm/transformers_utils/configs/dotsocr.py:12-31`` (DotsVisionConfig)
* decoder       -- Qwen2.5-1.5B shape, ``vllm/transformers_utils/configs/dotsocr.py:53-66``
                   and SURVEY.md Appendix A.2
* image maths   -- ``dots_ocr/utils/consts.py:1-3`` (reference)

Two presets exist: ``full()`` is the real architecture (what bench.py times),
``tiny()`` keeps head_dim 128 / patch 14 / merge 2 / GQA 6:1 so every kernel
path is exercised by CPU-sized oracle runs.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict


@dataclass(frozen=True)
class VisionConfig:
    embed_dim: int = 1536
    hidden_size: int = 1536          # merger output width == LLM hidden
    intermediate_size: int = 4224
    num_hidden_layers: int = 42
    num_attention_heads: int = 12
    num_channels: int = 3
    patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 1
    rms_norm_eps: float = 1e-5
    merger_ln_eps: float = 1e-6
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_attention_heads

    @property
    def patch_dim(self) -> int:
        return self.num_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def merge_dim(self) -> int:
        return self.embed_dim * self.spatial_merge_size ** 2


@dataclass(frozen=True)
class TextConfig:
    hidden_size: int = 1536
    intermediate_size: int = 8960
    num_hidden_layers: int = 28
    num_attention_heads: int = 12
    num_key_value_heads: int = 2
    head_dim: int = 128
    vocab_size: int = 151936
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    max_position_embeddings: int = 131072

    @property
    def q_dim(self) -> int:
        return self.num_attention_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_key_value_heads * self.head_dim


@dataclass(frozen=True)
class DotsConfig:
    vision: VisionConfig = field(default_factory=VisionConfig)
    text: TextConfig = field(default_factory=TextConfig)
    image_token_id: int = 151665
    video_token_id: int = 151656
    name: str = "dots.ocr"

    def to_dict(self) -> dict:
        return asdict(self)


def full() -> DotsConfig:
    """The real dots.ocr architecture (ViT 1.26 B + decoder 1.78 B)."""
    return DotsConfig()


def tiny() -> DotsConfig:
    """Same operator set at CPU-oracle size (2+2 layers)."""
    v = VisionConfig(embed_dim=256, hidden_size=768, intermediate_size=512,
                     num_hidden_layers=2, num_attention_heads=2)
    t = TextConfig(hidden_size=768, intermediate_size=1024, num_hidden_layers=2,
                   num_attention_heads=6, num_key_value_heads=1, head_dim=128,
                   vocab_size=2048, max_position_embeddings=8192)
    return DotsConfig(vision=v, text=t, image_token_id=2040, video_token_id=2041,
                      name="dots.ocr-tiny")


def small() -> DotsConfig:
    """Full widths, few layers: per-op shapes identical to ``full()``."""
    v = VisionConfig(num_hidden_layers=2)
    t = TextConfig(num_hidden_layers=2)
    return DotsConfig(vision=v, text=t, name="dots.ocr-small")


PRESETS = {"full": full, "tiny": tiny, "small": small}


class UnsupportedCheckpoint(ValueError):
    """The checkpoint's config.json asks for something the sm_100a kernels do not implement."""


def from_hf_dict(d: dict) -> DotsConfig:
    """DotsConfig from the ``config.json`` of a HF ``weights/DotsOCR`` directory (``dots_ocr/parser.py:67``;
    field names per ``vllm/transformers_utils/configs/dotsocr.py:12-66``: Qwen2Config at the top level,
    ``vision_config`` nested).  Missing keys take the published defaults; options that would change the
    arithmetic the kernels implement are refused instead of being ignored."""
    vd = dict(d.get("vision_config") or {})
    bad = []
    if vd.get("use_bias", False):
        bad.append("vision_config.use_bias=true (ViT linears are bias-free in the kernels)")
    if not vd.get("post_norm", True):
        bad.append("vision_config.post_norm=false")
    if vd.get("is_causal", False):
        bad.append("vision_config.is_causal=true")
    if d.get("tie_word_embeddings", False):
        bad.append("tie_word_embeddings=true (lm_head is a separate matrix)")
    if d.get("rope_scaling") not in (None, {}) and (d["rope_scaling"] or {}).get("rope_type", "default") != "default":
        bad.append(f"rope_scaling={d['rope_scaling']!r}")
    if isinstance(d.get("rope_parameters"), dict) and d["rope_parameters"].get("rope_type", "default") != "default":
        bad.append(f"rope_parameters={d['rope_parameters']!r}")
    if d.get("use_sliding_window", False):
        bad.append("use_sliding_window=true")
    if d.get("hidden_act", "silu") != "silu":
        bad.append(f"hidden_act={d.get('hidden_act')!r}")
    vkeys = ("embed_dim", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_channels",
             "patch_size", "spatial_merge_size", "temporal_patch_size", "rms_norm_eps")
    v = VisionConfig(**{k: vd[k] for k in vkeys if k in vd})
    heads = int(d.get("num_attention_heads", TextConfig.num_attention_heads))
    hidden = int(d.get("hidden_size", TextConfig.hidden_size))
    tkw = dict(hidden_size=hidden, num_attention_heads=heads, head_dim=int(d.get("head_dim") or hidden // heads))
    for k in ("intermediate_size", "num_hidden_layers", "num_key_value_heads", "vocab_size", "rms_norm_eps", "rope_theta",
              "max_position_embeddings"):
        if k in d:
            tkw[k] = d[k]
    if "rope_theta" not in d and isinstance(d.get("rope_parameters"), dict) and "rope_theta" in d["rope_parameters"]:
        tkw["rope_theta"] = d["rope_parameters"]["rope_theta"]      # transformers >= 5 spelling
    t = TextConfig(**tkw)
    if t.head_dim != 128 or v.head_dim != 128:
        bad.append(f"head_dim {t.head_dim} (text) / {v.head_dim} (vision): the attention kernels are built for 128")
    if v.hidden_size != t.hidden_size:
        bad.append(f"vision_config.hidden_size {v.hidden_size} != hidden_size {t.hidden_size}")
    if t.num_attention_heads % t.num_key_value_heads:
        bad.append("num_attention_heads is not a multiple of num_key_value_heads")
    if v.temporal_patch_size != 1 or v.spatial_merge_size != 2 or v.patch_size != 14 or v.num_channels != 3:
        bad.append("patch geometry other than 3x1x14x14 with 2x2 merge")
    if bad:
        raise UnsupportedCheckpoint("config.json is outside what the B200 path implements: " + "; ".join(bad))
    return DotsConfig(vision=v, text=t, image_token_id=int(d.get("image_token_id", 151665)),
                      video_token_id=int(d.get("video_token_id", 151656)), name=str(d.get("_name_or_path") or "dots.ocr"))


def from_hf_dir(path: str) -> DotsConfig:
    import json
    import os
    with open(os.path.join(path, "config.json"), "r", encoding="utf-8") as f:
        return from_hf_dict(json.load(f))
