"""Post-decode layout handling: map the model's bounding boxes between the original page and the
smart-resized image the model saw, and interpret the decoded response.

Mirrors the call surface of ``dots_ocr/utils/layout_utils.py`` (``pre_process_bboxes`` :115-144, ``post_process_cells``
:146-193, ``is_legal_bbox`` :195-200, ``post_process_output`` :202-228); behaviour is pinned against the reference functions
executed in the build container (``tests/golden/postprocess.json``).  A response that does not parse goes through
``output_cleaner.OutputCleaner`` (the reference's JSON repair) and comes back as text with ``filtered=True``.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence, Tuple

from .consts import MIN_PIXELS, MAX_PIXELS
from .image_utils import smart_resize

TEXT_ONLY_MODES = ("prompt_ocr", "prompt_table_html", "prompt_table_latex", "prompt_formula_latex")


def _model_size(input_width: int, input_height: int, min_pixels: Optional[int], max_pixels: Optional[int]) -> Tuple[int, int]:
    """(width, height) of the image the model really saw: the server side applies smart_resize once more."""
    h, w = smart_resize(input_height, input_width, min_pixels=min_pixels or MIN_PIXELS, max_pixels=max_pixels or MAX_PIXELS)
    return w, h


def _rescale(bbox: Sequence, sx: float, sy: float) -> List[int]:
    # the reference divides by the scale and truncates towards zero, coordinate by coordinate
    return [int(float(bbox[0]) / sx), int(float(bbox[1]) / sy), int(float(bbox[2]) / sx), int(float(bbox[3]) / sy)]


def pre_process_bboxes(origin_image, bboxes: List[List], input_width: int, input_height: int, factor: int = 28,
                       min_pixels: int = 3136, max_pixels: int = 11289600) -> List[List[int]]:
    """Original-page boxes -> coordinates in the image the model sees (grounding prompts)."""
    assert isinstance(bboxes, list) and len(bboxes) > 0 and isinstance(bboxes[0], list)
    ow, oh = origin_image.size
    mw, mh = _model_size(input_width, input_height, min_pixels, max_pixels)
    return [_rescale(b, ow / mw, oh / mh) for b in bboxes]


def post_process_cells(origin_image, cells: List[Dict], input_width: int, input_height: int, factor: int = 28,
                       min_pixels: int = 3136, max_pixels: int = 11289600) -> List[Dict]:
    """Model-space boxes of the decoded layout cells -> original-page coordinates (other keys are kept)."""
    assert isinstance(cells, list) and len(cells) > 0 and isinstance(cells[0], dict)
    ow, oh = origin_image.size
    mw, mh = _model_size(input_width, input_height, min_pixels, max_pixels)
    sx, sy = mw / ow, mh / oh
    out = []
    for cell in cells:
        c = dict(cell)
        c["bbox"] = _rescale(cell["bbox"], sx, sy)
        out.append(c)
    return out


def is_legal_bbox(cells: List[Dict]) -> bool:
    return all(c["bbox"][2] > c["bbox"][0] and c["bbox"][3] > c["bbox"][1] for c in cells)


def post_process_output(response, prompt_mode: str, origin_image, input_image, min_pixels=None, max_pixels=None):
    """Decoded text -> (cells in page coordinates, filtered=False), or (repaired text, True) when the layout JSON does not parse.
    Text-only prompt modes return the response unchanged (as the reference does)."""
    if prompt_mode in TEXT_ONLY_MODES:
        return response
    cells = response
    try:
        cells = json.loads(cells)
        return post_process_cells(origin_image, cells, input_image.width, input_image.height,
                                  min_pixels=min_pixels, max_pixels=max_pixels), False
    except Exception:       # noqa: BLE001 -- not JSON, not a list of cells, or a cell without a usable bbox
        pass
    from .output_cleaner import OutputCleaner
    repaired = OutputCleaner().clean_model_output(cells)
    if isinstance(repaired, list):
        repaired = "\n\n".join(cell["text"] for cell in repaired if "text" in cell)
    return repaired, True


# ------------------------------------------------------------------ layout overlay
# Category colours of the reference's overlay (dots_ocr/utils/layout_utils.py:14-28); anything else is drawn green.
LAYOUT_COLORS = {
    "Text": (0, 128, 0), "Footnote": (0, 128, 0), "Page-header": (0, 128, 0),
    "Picture": (255, 0, 255), "Caption": (255, 165, 0), "Section-header": (0, 255, 255), "Formula": (128, 128, 128),
    "Table": (255, 192, 203), "Title": (255, 0, 0), "List-item": (0, 0, 255), "Page-footer": (128, 0, 128),
    "Other": (165, 42, 42), "Unknown": (0, 0, 0),
}
FILL_OPACITY = 0.3          # layout_utils.py:90
LABEL_PX = 20               # label font size; the label "<reading order>_<category>" sits right of the box's top edge (:103-106)


def draw_layout_on_image(image, cells, resized_height=None, resized_width=None, fill_bbox=True, draw_bbox=True):
    """The page with every layout cell marked: a 30 % opaque fill (or a thin outline when ``fill_bbox`` is False) in the
    category colour and the label ``<index>_<category>`` at the box's top-right corner -- what the reference renders through
    a PyMuPDF page (``layout_utils.py:31-112``), done here with PIL compositing so it needs no PDF library.  Pixel-for-pixel
    equality with PyMuPDF's rasteriser is not a goal (fonts and anti-aliasing differ); geometry, colours and opacity are the
    reference's.  ``resized_*``: the cells are in the resized image's coordinates and are mapped back to ``image``'s."""
    from PIL import Image, ImageDraw, ImageFont
    base = image.convert("RGBA")
    W, H = base.size
    sx = (resized_width / W) if (resized_height and resized_width) else 1.0
    sy = (resized_height / H) if (resized_height and resized_width) else 1.0
    overlay = Image.new("RGBA", base.size, (0, 0, 0, 0))
    draw = ImageDraw.Draw(overlay)
    try:
        font = ImageFont.load_default(size=LABEL_PX)
    except TypeError:            # Pillow < 10.1: fixed-size bitmap font
        font = ImageFont.load_default()
    for order, cell in enumerate(cells):
        x0, y0, x1, y1 = (int(cell["bbox"][0] / sx), int(cell["bbox"][1] / sy), int(cell["bbox"][2] / sx), int(cell["bbox"][3] / sy))
        rgb = LAYOUT_COLORS.get(cell.get("category"), (0, 128, 0))
        if draw_bbox and x1 >= x0 and y1 >= y0:
            if fill_bbox:
                draw.rectangle([x0, y0, x1, y1], fill=rgb + (int(round(255 * FILL_OPACITY)),))
            else:
                draw.rectangle([x0, y0, x1, y1], outline=rgb + (255,), width=1)
        draw.text((x1, y0), f"{order}_{cell.get('category')}", fill=rgb + (255,), font=font)
    return Image.alpha_composite(base, overlay).convert("RGB")
