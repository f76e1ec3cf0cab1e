"""Host-side input contract of the hot path: PIL image + prompt -> the tensors ``generate`` takes.

Restates the stock ``Qwen2VLImageProcessor`` (fast/torchvision variant the reference selects with
``use_fast=True``, ``dots_ocr/parser.py:75``) for the dots.ocr settings: patch 14, merge 2,
temporal_patch_size 1, CLIP mean/std, bicubic + antialias resize on the uint8 tensor
(``transformers/models/qwen2_vl/image_processing_qwen2_vl.py:148-232``).
``tests/test_cpu_host.py::test_preprocess_image_equals_hf_processor`` checks bit-equality with the transformers class.

The chat template and tokenizer live in the HF checkpoint directory, which does not exist offline;
``SyntheticTokenizer`` is a byte-level stand-in used only to exercise the plumbing.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .utils.consts import MIN_PIXELS, MAX_PIXELS
from .utils.image_utils import smart_resize

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
IMAGE_PLACEHOLDER = "<|img|><|imgpad|><|endofimg|>"      # dots_ocr/model/inference.py:33


def to_rgb(pil_image):
    """RGBA is composited on white, everything else converted (reference image_utils.py:74-80)."""
    from PIL import Image
    if pil_image.mode == "RGBA":
        bg = Image.new("RGB", pil_image.size, (255, 255, 255))
        bg.paste(pil_image, mask=pil_image.split()[3])
        return bg
    return pil_image.convert("RGB")


def preprocess_image(image, min_pixels: Optional[int] = None, max_pixels: Optional[int] = None,
                     patch: int = 14, merge: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """PIL image (or uint8 HWC array) -> (pixel_values [gh*gw, 3*patch*patch] fp32, grid_thw [1, 3] int64).
    Rows are in 2x2 merge-block order; each row is channel-major (c, py, px)."""
    import torchvision.transforms.v2.functional as tvF
    from torchvision.transforms import InterpolationMode
    if not isinstance(image, (np.ndarray, torch.Tensor)):
        image = np.asarray(to_rgb(image))
    img = torch.from_numpy(np.array(image, copy=True))
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3, "expect uint8 HWC RGB"
    img = img.permute(2, 0, 1).contiguous()                                 # CHW
    H, W = img.shape[1:]
    rh, rw = smart_resize(H, W, factor=patch * merge, min_pixels=min_pixels or MIN_PIXELS,
                          max_pixels=max_pixels or MAX_PIXELS)
    if (rh, rw) != (H, W):
        img = tvF.resize(img, [rh, rw], interpolation=InterpolationMode.BICUBIC, antialias=True)
    x = img.to(torch.float32)
    # the fast processor fuses rescale into the normalisation: (x - 255*mean) / (255*std)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32) * 255.0
    std = torch.tensor(CLIP_STD, dtype=torch.float32) * 255.0
    x = (x - mean[:, None, None]) / std[:, None, None]
    gh, gw = rh // patch, rw // patch
    x = x.view(3, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.permute(1, 4, 2, 5, 0, 3, 6).reshape(gh * gw, 3 * patch * patch)   # (bh, bw, ih, iw | c, py, px)
    return x.contiguous(), torch.tensor([[1, gh, gw]], dtype=torch.int64)


def preprocess_image_u8(image, min_pixels: Optional[int] = None, max_pixels: Optional[int] = None, patch: int = 14,
                        merge: int = 2) -> torch.Tensor:
    """Host half of the GPU pre-processing path: PIL image -> smart-resized uint8 HWC tensor (the bicubic + antialias resize stays
    on the CPU so that it is the same torchvision resize the stock processor runs); rescale / normalise / patchify happen in
    dots_patchify_u8 on the device."""
    import torchvision.transforms.v2.functional as tvF
    from torchvision.transforms import InterpolationMode
    if not isinstance(image, (np.ndarray, torch.Tensor)):
        image = np.asarray(to_rgb(image))
    img = torch.from_numpy(np.array(image, copy=True))
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3, "expect uint8 HWC RGB"
    H, W = img.shape[:2]
    rh, rw = smart_resize(H, W, factor=patch * merge, min_pixels=min_pixels or MIN_PIXELS, max_pixels=max_pixels or MAX_PIXELS)
    if (rh, rw) != (H, W):
        chw = tvF.resize(img.permute(2, 0, 1).contiguous(), [rh, rw], interpolation=InterpolationMode.BICUBIC, antialias=True)
        img = chw.permute(1, 2, 0)
    return img.contiguous()


def page_to_u8(image) -> torch.Tensor:
    """PIL image (any mode) or uint8 HWC array -> uint8 HWC RGB tensor at its ORIGINAL size (alpha composited on white like the
    reference, image_utils.py:74-80).  Everything after this -- resize, rescale, normalise, patchify -- runs on the GPU."""
    if not isinstance(image, (np.ndarray, torch.Tensor)):
        image = np.asarray(to_rgb(image))
    img = torch.from_numpy(np.array(image, copy=True)) if not isinstance(image, torch.Tensor) else image
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3, "expect uint8 HWC RGB"
    return img.contiguous()


def model_image_tokens(h: int, w: int, min_pixels: Optional[int] = None, max_pixels: Optional[int] = None, patch: int = 14, merge: int = 2) -> int:
    """<|imgpad|> slots a page of h x w pixels occupies after smart_resize (image tokens = patches / merge^2)."""
    rh, rw = smart_resize(h, w, factor=patch * merge, min_pixels=min_pixels or MIN_PIXELS, max_pixels=max_pixels or MAX_PIXELS)
    return (rh // patch) * (rw // patch) // (merge * merge)


class SyntheticTokenizer:
    """Byte-level stand-in (ids 0..255 = bytes) with the three image specials mapped onto the config's
    reserved ids.  NOT the dots.ocr tokenizer: decoded text is meaningless with synthetic weights."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.image_token_id = cfg.image_token_id
        self.img_start_id = cfg.image_token_id - 2
        self.img_end_id = cfg.image_token_id - 1
        self.user_id, self.end_user_id, self.assistant_id = (cfg.image_token_id - 5, cfg.image_token_id - 4,
                                                             cfg.image_token_id - 3)
        self.eos_token_id = None
        self.pad_token_id = 0

    def encode_chat(self, prompt: str, n_image_tokens: int) -> List[int]:
        """<|user|><|img|><|imgpad|>*n<|endofimg|>{prompt}<|endofuser|><|assistant|>."""
        ids = [self.user_id, self.img_start_id] + [self.image_token_id] * n_image_tokens + [self.img_end_id]
        ids += list(prompt.encode("utf-8"))
        ids += [self.end_user_id, self.assistant_id]
        return ids

    def decode(self, ids: Sequence[int]) -> str:
        return bytes(int(i) for i in ids if 0 <= int(i) < 256).decode("utf-8", errors="replace")


class HFTokenizer:
    """The checkpoint's own tokenizer behind the interface ``PageRunner`` uses (``encode_chat`` / ``decode`` /
    ``eos_token_id`` / ``pad_token_id`` / ``stop_ids``).

    The prompt text follows the reference's two call sites: the chat template applied to a user turn whose content is
    ``<|img|><|imgpad|><|endofimg|>{prompt}`` (``dots_ocr/model/inference.py:25-36`` with vLLM's
    ``--chat-template-content-format string``; the HF processor renders the same string, ``parser.py:79-98``), after
    which the single ``<|imgpad|>`` is widened to one id per merged image token, as the processor does.  When the
    directory ships no chat template, the layout documented in SURVEY.md Appendix C is used:
    ``<|user|>…<|endofuser|><|assistant|>``."""

    IMG, PAD, END = "<|img|>", "<|imgpad|>", "<|endofimg|>"

    def __init__(self, path_or_tokenizer, image_token_id: Optional[int] = None, generation_config: Optional[dict] = None):
        if isinstance(path_or_tokenizer, str):
            from transformers import AutoTokenizer
            # the published checkpoint ships its own configuration / processor classes and the reference loads it with
            # trust_remote_code=True (dots_ocr/parser.py:68-75); without it transformers stops at an interactive prompt
            tok = AutoTokenizer.from_pretrained(path_or_tokenizer, trust_remote_code=True)
            if generation_config is None:
                import json
                import os
                g = os.path.join(path_or_tokenizer, "generation_config.json")
                if os.path.isfile(g):
                    with open(g, "r", encoding="utf-8") as f:
                        generation_config = json.load(f)
        else:
            tok = path_or_tokenizer
        self.tok = tok
        pad = tok.convert_tokens_to_ids(self.PAD)
        if pad is None or pad == getattr(tok, "unk_token_id", None):
            raise ValueError("tokenizer has no <|imgpad|> token: not a dots.ocr tokenizer")
        if image_token_id is not None and int(image_token_id) != int(pad):
            raise ValueError(f"config image_token_id {image_token_id} != tokenizer id of <|imgpad|> {pad}")
        self.image_token_id = int(pad)
        # HF generate's stop set: generation_config.json's eos_token_id (int or list), else the tokenizer's eos
        eos = (generation_config or {}).get("eos_token_id", tok.eos_token_id)
        ids = [] if eos is None else ([int(e) for e in eos] if isinstance(eos, (list, tuple)) else [int(eos)])
        self.stop_ids = tuple(dict.fromkeys(ids))
        self.eos_token_id = self.stop_ids[0] if self.stop_ids else None
        p = (generation_config or {}).get("pad_token_id", tok.pad_token_id)
        self.pad_token_id = int(p) if p is not None else (self.eos_token_id if self.eos_token_id is not None else 0)
        self._cache = {}

    def render(self, prompt: str) -> str:
        content = f"{self.IMG}{self.PAD}{self.END}{prompt}"
        if getattr(self.tok, "chat_template", None):
            return self.tok.apply_chat_template([{"role": "user", "content": content}], tokenize=False,
                                                add_generation_prompt=True)
        return f"<|user|>{content}<|endofuser|><|assistant|>"

    def encode_chat(self, prompt: str, n_image_tokens: int) -> List[int]:
        halves = self._cache.get(prompt)
        if halves is None:
            text = self.render(prompt)
            if text.count(self.PAD) != 1:
                raise ValueError("the rendered prompt must contain exactly one <|imgpad|> (one image per page)")
            before, after = text.split(self.PAD)
            # <|imgpad|> is a special token, so tokenising either side of it separately equals tokenising the whole
            halves = (self.tok.encode(before, add_special_tokens=False), self.tok.encode(after, add_special_tokens=False))
            if len(self._cache) < 64:
                self._cache[prompt] = halves
        return list(halves[0]) + [self.image_token_id] * int(n_image_tokens) + list(halves[1])

    def decode(self, ids: Sequence[int]) -> str:
        """``processor.batch_decode(..., skip_special_tokens=True, clean_up_tokenization_spaces=False)``
        (parser.py:114-116), cut at the first stop id."""
        ids = [int(i) for i in ids]
        for k, i in enumerate(ids):
            if i in self.stop_ids:
                ids = ids[:k]
                break
        return self.tok.decode(ids, skip_special_tokens=True, clean_up_tokenization_spaces=False)


def build_text_inputs(tokenizer: SyntheticTokenizer, n_image_tokens: Sequence[int], prompts: Sequence[str]):
    """Left-padded input_ids + attention_mask for prompts whose images contribute `n_image_tokens[i]` <|imgpad|> slots."""
    rows = [tokenizer.encode_chat(prompt, n) for prompt, n in zip(prompts, n_image_tokens)]
    T = max(len(r) for r in rows)
    ids = torch.full((len(rows), T), tokenizer.pad_token_id, dtype=torch.int64)
    mask = torch.zeros((len(rows), T), dtype=torch.int64)
    for i, r in enumerate(rows):
        ids[i, T - len(r):] = torch.tensor(r)
        mask[i, T - len(r):] = 1
    return dict(input_ids=ids, attention_mask=mask)


def build_inputs(tokenizer: SyntheticTokenizer, images: Sequence, prompts: Sequence[str], min_pixels=None, max_pixels=None,
                 merge: int = 2):
    """What ``processor(text=[...], images=[...], padding=True, return_tensors="pt")`` returns
    (parser.py:99-105): left-padded input_ids + attention_mask, concatenated pixel_values, grid_thw."""
    pvs, grids, rows = [], [], []
    for img, prompt in zip(images, prompts):
        pv, g = preprocess_image(img, min_pixels, max_pixels, merge=merge)
        pvs.append(pv)
        grids.append(g)
        rows.append(tokenizer.encode_chat(prompt, pv.shape[0] // (merge * merge)))
    T = max(len(r) for r in rows)
    ids = torch.full((len(rows), T), tokenizer.pad_token_id, dtype=torch.int64)
    mask = torch.zeros((len(rows), T), dtype=torch.int64)
    for i, r in enumerate(rows):
        ids[i, T - len(r):] = torch.tensor(r)
        mask[i, T - len(r):] = 1
    return dict(input_ids=ids, attention_mask=mask, pixel_values=torch.cat(pvs), image_grid_thw=torch.cat(grids))
