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


def _open_image(source):
    """PIL image from what the reference's ``fetch_image`` accepts (``dots_ocr/utils/image_utils.py:84-111``): a PIL image,
    an http(s) URL, a ``file://`` URL, a ``data:image...;base64,`` URL or a local path."""
    import io
    from PIL import Image
    if isinstance(source, Image.Image):
        return source
    if not isinstance(source, str):
        raise ValueError(f"Unrecognized image input, support local path, http url, base64 and PIL.Image, got {source!r}")
    payload = None
    if source.startswith(("http://", "https://")):
        import urllib.request
        with urllib.request.urlopen(source, timeout=60) as resp:
            payload = resp.read()
    elif source.startswith("data:image"):
        import base64
        if "base64," not in source:
            raise ValueError("Unrecognized image input: a data: URL must carry base64 data")
        payload = base64.b64decode(source.split("base64,", 1)[1])
    if payload is not None:
        img = Image.open(io.BytesIO(payload))
        img.load()                       # decode now: the buffer does not outlive this call
        return img
    return Image.open(source[7:] if source.startswith("file://") else source)


def fetch_image(image, min_pixels=None, max_pixels=None, resized_height=None, resized_width=None):
    """Load + RGB-convert (alpha composited on white) + resize as the reference does before the model call
    (``image_utils.py:84-138``): an explicit ``resized_height/width`` is snapped to the 28-px grid; otherwise, when a pixel
    budget is given, the image's own size goes through ``smart_resize`` with that budget.  No budget: size unchanged."""
    from ..processing import to_rgb
    assert image is not None, f"image not found, maybe input format error: {image}"
    img = to_rgb(_open_image(image))
    target = None
    if resized_height and resized_width:
        target = smart_resize(resized_height, resized_width, factor=IMAGE_FACTOR)
    elif min_pixels or max_pixels:
        target = smart_resize(img.height, img.width, factor=IMAGE_FACTOR, min_pixels=min_pixels or MIN_PIXELS,
                              max_pixels=max_pixels or MAX_PIXELS)
    if target is not None:
        img = img.resize((target[1], target[0]))
    return img
