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
