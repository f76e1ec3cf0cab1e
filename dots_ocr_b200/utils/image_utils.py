"""Resize arithmetic that fixes every downstream tensor shape on the hot path.

``smart_resize`` restates ``dots_ocr/utils/image_utils.py:29-63`` of the
reference (itself derived from Qwen2.5-VL's vision_process.py).  Known answers
taken by executing the reference function are committed in
``tests/golden/smart_resize.json`` (see ``tests/golden/make_smart_resize_golden.py``).
"""
from __future__ import annotations

import math
from typing import Tuple

from .consts import IMAGE_FACTOR, MIN_PIXELS, MAX_PIXELS


def round_by_factor(number: float, factor: int) -> int:
    return round(number / factor) * factor


def ceil_by_factor(number: float, factor: int) -> int:
    return math.ceil(number / factor) * factor


def floor_by_factor(number: float, factor: int) -> int:
    return math.floor(number / factor) * factor


def smart_resize(height: int, width: int, factor: int = IMAGE_FACTOR,
                 min_pixels: int = MIN_PIXELS, max_pixels: int = MAX_PIXELS) -> Tuple[int, int]:
    """Target (H, W): both multiples of ``factor``, area within
    [min_pixels, max_pixels], aspect ratio kept as closely as possible.
    Raises ValueError when the aspect ratio exceeds 200 (reference :45-48)."""
    long_side, short_side = max(height, width), min(height, width)
    if long_side / short_side > 200:
        raise ValueError(
            f"absolute aspect ratio must be smaller than 200, got {long_side / short_side}")
    h_bar = max(factor, round_by_factor(height, factor))
    w_bar = max(factor, round_by_factor(width, factor))
    area = h_bar * w_bar
    if area > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, floor_by_factor(height / beta, factor))
        w_bar = max(factor, floor_by_factor(width / beta, factor))
    elif area < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = ceil_by_factor(height * beta, factor)
        w_bar = ceil_by_factor(width * beta, factor)
        if h_bar * w_bar > max_pixels:      # max_pixels wins: it bounds the token count
            beta = math.sqrt((h_bar * w_bar) / max_pixels)
            h_bar = max(factor, floor_by_factor(h_bar / beta, factor))
            w_bar = max(factor, floor_by_factor(w_bar / beta, factor))
    return h_bar, w_bar


def vit_grid(height: int, width: int, patch: int = 14, **kw) -> Tuple[int, int]:
    """(grid_h, grid_w) in 14-pixel patches after ``smart_resize``."""
    h, w = smart_resize(height, width, **kw)
    return h // patch, w // patch


def token_counts(height: int, width: int, patch: int = 14, merge: int = 2, **kw) -> Tuple[int, int]:
    """(ViT tokens, LLM image tokens) for an image of the given size."""
    gh, gw = vit_grid(height, width, patch, **kw)
    return gh * gw, (gh * gw) // (merge * merge)


def PILimage_to_base64(image, format: str = "PNG") -> str:
    """``data:image/<fmt>;base64,...`` URL of a PIL image: what the reference's HTTP client puts in ``image_url``
    (``dots_ocr/utils/image_utils.py:64-68``) and what ``dots_ocr_b200.server`` accepts."""
    import base64
    import io
    with io.BytesIO() as buf:
        image.save(buf, format=format)
        payload = base64.b64encode(buf.getvalue()).decode("ascii")
    return "data:image/" + format.lower() + ";base64," + payload
