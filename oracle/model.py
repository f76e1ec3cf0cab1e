"""ORACLE (test infrastructure): whole-page forward = vision oracle + HF Qwen2ForCausalLM.

Call order follows the reference's HF path (``dots_ocr/parser.py:99-116``,
``demo/demo_hf.py:33-50``): embed ids, overwrite ``<|imgpad|>`` rows with the ViT output
(``masked_scatter``), run the stock Qwen2 decoder greedily through
``GenerationMixin.generate`` (``transformers/generation/utils.py:2658-2805``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .vision import VisionOracle


def build_qwen2(tcfg, ckpt: Dict[str, torch.Tensor], dtype, device, attn_impl: str = "sdpa"):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    hf = Qwen2Config(vocab_size=tcfg.vocab_size, hidden_size=tcfg.hidden_size,
                     intermediate_size=tcfg.intermediate_size, num_hidden_layers=tcfg.num_hidden_layers,
                     num_attention_heads=tcfg.num_attention_heads, num_key_value_heads=tcfg.num_key_value_heads,
                     max_position_embeddings=tcfg.max_position_embeddings, rms_norm_eps=tcfg.rms_norm_eps,
                     rope_theta=tcfg.rope_theta, tie_word_embeddings=False, attn_implementation=attn_impl)
    with torch.device("meta"):
        m = Qwen2ForCausalLM(hf)
    m = m.to_empty(device=device).to(dtype)
    sd = {k: v.to(device=device, dtype=dtype) for k, v in ckpt.items() if not k.startswith("vision_tower.")}
    missing, unexpected = m.load_state_dict(sd, strict=False, assign=True)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    # rotary inv_freq is a non-persistent fp32 buffer: to_empty() left it uninitialised and .to(dtype)
    # would round it; rebuild it exactly as Qwen2RotaryEmbedding.__init__ does (modeling_qwen2.py:54-68).
    rot = m.model.rotary_emb
    inv, scaling = rot.compute_default_rope_parameters(rot.config, device)
    rot.inv_freq = inv.to(device=device, dtype=torch.float32)
    rot.original_inv_freq = rot.inv_freq.clone()
    rot.attention_scaling = scaling
    return m.eval()


class DotsOracle:
    def __init__(self, cfg, ckpt: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu",
                 attn_impl: str = "sdpa"):
        self.cfg = cfg
        self.dtype = dtype
        self.device = torch.device(device)
        self.vision = VisionOracle(cfg.vision, ckpt, dtype=dtype, device=device)
        self.llm = build_qwen2(cfg.text, ckpt, dtype, self.device, attn_impl)

    @torch.no_grad()
    def inputs_embeds(self, input_ids, pixel_values=None, image_grid_thw=None):
        ids = input_ids.to(self.device)
        emb = self.llm.model.embed_tokens(ids)
        if pixel_values is not None:
            img = self.vision.forward(pixel_values, image_grid_thw).to(emb.dtype)
            mask = (ids == self.cfg.image_token_id)
            assert int(mask.sum()) == img.shape[0], (int(mask.sum()), img.shape)
            emb = emb.masked_scatter(mask.unsqueeze(-1).expand_as(emb), img)
        return emb

    @torch.no_grad()
    def generate(self, input_ids, attention_mask=None, pixel_values=None, image_grid_thw=None,
                 max_new_tokens: int = 16, eos_token_id: Optional[int] = None):
        """Returns ids [B, T+N] including the prompt, like ``model.generate(**inputs)`` (parser.py:110)."""
        ids = input_ids.to(self.device)
        if attention_mask is None:
            attention_mask = torch.ones_like(ids)
        emb = self.inputs_embeds(ids, pixel_values, image_grid_thw)
        new = self.llm.generate(inputs_embeds=emb, attention_mask=attention_mask.to(self.device),
                                max_new_tokens=max_new_tokens, do_sample=False,
                                eos_token_id=eos_token_id, pad_token_id=0)
        return torch.cat([ids, new], dim=1)

    @torch.no_grad()
    def teacher_forced_logits(self, input_ids, new_ids, pixel_values=None, image_grid_thw=None):
        """fp32 logits [B, N, V] that predict ``new_ids[:, j]`` given prompt + new_ids[:, :j]
        (one full-sequence forward; causal masking makes it equal to step-by-step decode)."""
        ids = input_ids.to(self.device)
        new_ids = new_ids.to(self.device)
        emb = self.inputs_embeds(ids, pixel_values, image_grid_thw)
        if new_ids.shape[1] > 1:
            emb = torch.cat([emb, self.llm.model.embed_tokens(new_ids[:, :-1])], dim=1)
        hid = self.llm.model(inputs_embeds=emb).last_hidden_state
        T = ids.shape[1]
        return self.llm.lm_head(hid[:, T - 1:, :]).float()
