#!/usr/bin/env python
"""Fabricate a complete HF-style ``weights/DotsOCR`` directory from the seeded synthetic checkpoint.

No dots.ocr weights or tokenizer exist offline.  This writes everything a loader looks for --
``config.json`` (DotsOCRConfig layout, vllm/transformers_utils/configs/dotsocr.py:12-66), sharded ``*.safetensors`` with the
HF tensor names + ``model.safetensors.index.json``, a byte-level tokenizer whose special tokens sit on the ids the config
names (``<|imgpad|>`` = image_token_id), ``generation_config.json``, ``preprocessor_config.json`` (Qwen2VLImageProcessor,
the dots.ocr pixel limits) and a chat template in the documented layout -- so that

  * ``PageRunner.from_checkpoint(dir)`` exercises the real-checkpoint path of this package end to end, and
  * the in-image vLLM ``DotsOCRForCausalLM`` can be pointed at the SAME parameters for an on-box comparison.

    python tools/make_checkpoint_dir.py --preset tiny --out /tmp/dots_tiny            # seconds
    python tools/make_checkpoint_dir.py --preset full --flavour peaked --out weights/DotsOCR   # ~6 GB of bf16

The text the tokenizer produces is meaningless (byte-level, no merges); only shapes, ids and plumbing are real.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHAT_TEMPLATE = ("{% for m in messages %}"
                 "{% if m['role'] == 'system' %}<|system|>{{ m['content'] }}<|endofsystem|>"
                 "{% elif m['role'] == 'user' %}<|user|>{{ m['content'] }}<|endofuser|>"
                 "{% else %}<|assistant|>{{ m['content'] }}<|endofassistant|>{% endif %}"
                 "{% endfor %}{% if add_generation_prompt %}<|assistant|>{% endif %}")


def special_token_ids(cfg) -> dict:
    """Ids of the special tokens: the image trio and the turn markers sit where ``processing.SyntheticTokenizer`` puts them
    (image_token_id-5 .. image_token_id), the rest on free ids above."""
    I, vid, V = cfg.image_token_id, cfg.video_token_id, cfg.text.vocab_size
    ids = {"<|user|>": I - 5, "<|endofuser|>": I - 4, "<|assistant|>": I - 3, "<|img|>": I - 2, "<|endofimg|>": I - 1,
           "<|imgpad|>": I, "<|video_pad|>": vid}
    free = (i for i in range(I + 1, V) if i != vid)
    for name in ("<|endofassistant|>", "<|endoftext|>", "<|system|>", "<|endofsystem|>"):
        ids[name] = next(free)
    assert len(set(ids.values())) == len(ids) and min(ids.values()) >= 256, "vocabulary too small for the special tokens"
    return ids


def byte_chars() -> list:
    """The printable stand-in character of every byte value under the byte-level pre-tokenizer (the GPT-2 table: printable
    Latin-1 bytes map to themselves, the remaining 68 to U+0100 onwards in byte order)."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return [table[b] for b in range(256)]


def hf_config_dict(cfg) -> dict:
    t, v = cfg.text, cfg.vision
    sp = special_token_ids(cfg)
    return dict(
        architectures=["DotsOCRForCausalLM"], model_type="dots_ocr", torch_dtype="bfloat16",
        # like the published checkpoint, the directory carries its own configuration class (trust_remote_code): transformers and
        # vLLM 0.22 know no built-in `dots_ocr` model type; the class re-exports vLLM's DotsOCRConfig when vLLM is importable
        auto_map={"AutoConfig": "configuration_dots.DotsOCRConfig"},
        hidden_size=t.hidden_size, intermediate_size=t.intermediate_size, num_hidden_layers=t.num_hidden_layers,
        num_attention_heads=t.num_attention_heads, num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
        rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta, max_position_embeddings=t.max_position_embeddings,
        hidden_act="silu", tie_word_embeddings=False, attention_dropout=0.0, use_sliding_window=False, use_cache=True,
        image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
        bos_token_id=None, eos_token_id=sp["<|endoftext|>"], pad_token_id=sp["<|endoftext|>"],
        vision_config=dict(model_type="dots_vit", embed_dim=v.embed_dim, hidden_size=v.hidden_size,
                           intermediate_size=v.intermediate_size, num_hidden_layers=v.num_hidden_layers,
                           num_attention_heads=v.num_attention_heads, num_channels=v.num_channels, patch_size=v.patch_size,
                           spatial_merge_size=v.spatial_merge_size, temporal_patch_size=v.temporal_patch_size,
                           rms_norm_eps=v.rms_norm_eps, use_bias=False, attn_implementation="flash_attention_2",
                           initializer_range=0.02, init_merger_std=0.02, is_causal=False, post_norm=True,
                           gradient_checkpointing=False))


def write_tokenizer(cfg, out: str):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    V = cfg.text.vocab_size
    sp = special_token_ids(cfg)
    by_id = {b: c for b, c in enumerate(byte_chars())}       # id b = byte b, as in processing.SyntheticTokenizer
    assert len(by_id) == 256 and set(by_id.values()) == set(pre_tokenizers.ByteLevel.alphabet())
    for name, i in sp.items():
        by_id[i] = name
    for i in range(256, V):
        by_id.setdefault(i, f"<|fill_{i}|>")                 # never produced from text: the BPE model has no merges
    vocab = {tok: i for i, tok in by_id.items()}
    assert len(vocab) == V
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|endoftext|>", pad_token="<|endoftext|>",
                                   additional_special_tokens=[n for n in sp if n != "<|endoftext|>"])
    fast.chat_template = CHAT_TEMPLATE
    fast.save_pretrained(out)
    for name, i in sp.items():
        assert fast.convert_tokens_to_ids(name) == i, (name, i, fast.convert_tokens_to_ids(name))
    return fast


REMOTE_CONFIG = '''"""Configuration class shipped with the checkpoint directory (config.json: auto_map, trust_remote_code=True)."""
try:
    from vllm.transformers_utils.configs.dotsocr import DotsOCRConfig, DotsVisionConfig      # noqa: F401
except Exception:                                                                            # no vLLM: a plain Qwen2 config with the extra keys
    from transformers import Qwen2Config

    class DotsOCRConfig(Qwen2Config):
        model_type = "dots_ocr"

        def __init__(self, image_token_id=151665, video_token_id=151656, vision_config=None, **kw):
            super().__init__(**kw)
            self.image_token_id, self.video_token_id, self.vision_config = image_token_id, video_token_id, vision_config or {}
'''


def write_dir(cfg, out: str, seed: int = 0, flavour: str = "random", shards: int = 2) -> dict:
    from dots_ocr_b200 import weights as W
    os.makedirs(out, exist_ok=True)
    sp = special_token_ids(cfg)
    with open(os.path.join(out, "config.json"), "w") as f:
        json.dump(hf_config_dict(cfg), f, indent=2)
    with open(os.path.join(out, "configuration_dots.py"), "w") as f:
        f.write(REMOTE_CONFIG)
    with open(os.path.join(out, "generation_config.json"), "w") as f:
        json.dump({"do_sample": False, "eos_token_id": [sp["<|endoftext|>"], sp["<|endofassistant|>"]],
                   "pad_token_id": sp["<|endoftext|>"], "max_new_tokens": 24000}, f, indent=2)
    write_tokenizer(cfg, out)
    from transformers import Qwen2VLImageProcessor
    from dots_ocr_b200.utils.consts import MAX_PIXELS, MIN_PIXELS
    Qwen2VLImageProcessor(min_pixels=MIN_PIXELS, max_pixels=MAX_PIXELS, patch_size=cfg.vision.patch_size,
                          merge_size=cfg.vision.spatial_merge_size,
                          temporal_patch_size=cfg.vision.temporal_patch_size).save_pretrained(out)
    ckpt = W.make_synthetic_checkpoint(cfg, seed, flavour)
    W.save_safetensors_dir(ckpt, out, shards=shards)
    # the index file HF / vLLM loaders read for sharded checkpoints
    names = list(ckpt)
    per = (len(names) + shards - 1) // shards
    weight_map = {k: f"model-{i // per + 1:05d}-of-{shards:05d}.safetensors" for i, k in enumerate(names)}
    total = sum(v.numel() * v.element_size() for v in ckpt.values())
    with open(os.path.join(out, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f)
    return {"tensors": len(ckpt), "bytes": total, "special_tokens": sp}


def main() -> None:
    from dots_ocr_b200 import config as C
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--preset", default="tiny", choices=sorted(C.PRESETS))
    ap.add_argument("--flavour", default="random", choices=("random", "peaked"))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--shards", type=int, default=2)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    info = write_dir(C.PRESETS[a.preset](), a.out, a.seed, a.flavour, a.shards)
    print(json.dumps({"out": a.out, "preset": a.preset, "tensors": info["tensors"], "bytes": info["bytes"]}))


if __name__ == "__main__":
    main()
