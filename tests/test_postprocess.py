"""Post-decode CPU pipeline against golden vectors produced by EXECUTING the reference's own functions
(tests/golden/make_postprocess_golden.py; dots_ocr/utils/layout_utils.py:115-228, format_transformer.py:10-206)."""
import json
import os

from PIL import Image

from dots_ocr_b200.utils import format_transformer as F
from dots_ocr_b200.utils import layout_utils as L

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "postprocess.json")))


def test_post_process_cells_and_legality():
    assert len(G["cells"]) >= 50
    for c in G["cells"]:
        img = Image.new("RGB", tuple(c["origin"]))
        got = L.post_process_cells(img, c["cells"], c["input"][0], c["input"][1], **c["kw"])
        assert got == c["out"], c
        assert L.is_legal_bbox(c["cells"]) == c["legal"]
        assert c["cells"][0]["bbox"] is not got[0]["bbox"]            # inputs are not modified in place


def test_pre_process_bboxes():
    for c in G["bboxes"]:
        img = Image.new("RGB", tuple(c["origin"]))
        assert L.pre_process_bboxes(img, c["bboxes"], c["input"][0], c["input"][1], **c["kw"]) == c["out"], c


def test_formula_and_text_helpers():
    for c in G["formula"]:
        assert F.get_formula_in_markdown(c["in"]) == c["out"], c
    for c in G["has_latex"]:
        assert F.has_latex_markdown(c["in"]) == c["out"], c
    for c in G["preamble"]:
        assert F.clean_latex_preamble(c["in"]) == c["out"], c
    for c in G["clean_text"]:
        assert F.clean_text(c["in"]) == c["out"], c
    for c in G["fix"]:
        assert F.fix_streamlit_formulas(c["in"]) == c["out"], c


def test_layoutjson2md_including_picture_crops():
    for c in G["md"]:
        img = Image.new("RGB", tuple(c["size"]), tuple(c["color"]))
        assert F.layoutjson2md(img, c["cells"], no_page_hf=c["no_page_hf"]) == c["out"]


def test_post_process_output_modes():
    page, seen = Image.new("RGB", (1700, 2250)), Image.new("RGB", (1708, 2240))
    for c in G["output"]:
        r = L.post_process_output(c["response"], c["mode"], page, seen)
        assert isinstance(r, tuple) == c["tuple"]
        assert (list(r) if isinstance(r, tuple) else r) == c["out"], c
    # a response that is not JSON goes through the OutputCleaner: nothing recoverable -> empty text, flagged
    assert L.post_process_output("not json", "prompt_layout_all_en", page, seen) == ("", True)


def test_output_cleaner_against_the_reference_class():
    from dots_ocr_b200.utils.output_cleaner import OutputCleaner
    assert len(G["cleaner"]) >= 25
    for c in G["cleaner"]:
        assert OutputCleaner().clean_model_output(c["in"]) == c["out"], (str(c["in"])[:120], c["out"])


def test_post_process_output_failure_path():
    page, seen = Image.new("RGB", (1700, 2250)), Image.new("RGB", (1708, 2240))
    for c in G["output_fail"]:
        assert list(L.post_process_output(c["response"], "prompt_layout_all_en", page, seen)) == c["out"], c["response"][:120]


def test_layout_overlay_geometry_colours_and_opacity():
    """draw_layout_on_image: 30 % fill in the category colour inside the box, untouched outside, a label right of the
    top edge, outline-only mode, and cells given in resized coordinates mapped back to the page."""
    from PIL import Image
    from dots_ocr_b200.utils.layout_utils import LAYOUT_COLORS, draw_layout_on_image
    page = Image.new("RGB", (200, 100), "white")
    cells = [{"bbox": [10, 10, 60, 40], "category": "Text"}, {"bbox": [100, 50, 150, 90], "category": "Title"},
             {"bbox": [5, 60, 40, 95], "category": "NoSuchCategory"}]
    out = draw_layout_on_image(page, cells)
    assert out.size == page.size and out.mode == "RGB" and page.getpixel((30, 20)) == (255, 255, 255)      # input untouched

    def blend(rgb):
        a = round(255 * 0.3) / 255
        return tuple(int(round(255 * (1 - a) + c * a)) for c in rgb)
    for xy, cat in (((30, 25), "Text"), ((125, 87), "Title")):      # (125, 87): below the third cell's label
        got, want = out.getpixel(xy), blend(LAYOUT_COLORS[cat])
        assert all(abs(g - w) <= 1 for g, w in zip(got, want)), (cat, got, want)
    got = out.getpixel((20, 80))
    assert all(abs(g - w) <= 1 for g, w in zip(got, blend((0, 128, 0))))                 # unknown category: green
    assert out.getpixel((80, 5)) == (255, 255, 255) and out.getpixel((199, 99)) == (255, 255, 255)
    label = out.crop((61, 10, 140, 36))                                                   # "0_Text" right of the first box
    assert any(b != 255 for b in label.tobytes())

    line = draw_layout_on_image(page, cells[:1], fill_bbox=False)
    assert line.getpixel((10, 25)) == LAYOUT_COLORS["Text"] and line.getpixel((30, 25)) == (255, 255, 255)
    nothing = draw_layout_on_image(page, cells[:1], draw_bbox=False)
    assert nothing.getpixel((30, 25)) == (255, 255, 255)
    # cells in the coordinates of a 400x200 resize of the page
    big = draw_layout_on_image(page, [{"bbox": [20, 20, 120, 80], "category": "Table"}], resized_height=200, resized_width=400)
    got, want = big.getpixel((30, 25)), blend(LAYOUT_COLORS["Table"])
    assert all(abs(g - w) <= 1 for g, w in zip(got, want)) and big.getpixel((30, 60)) == (255, 255, 255)


def _fetch_inputs(tmpdir):
    """The inputs tests/golden/make_postprocess_golden.py:fetch_inputs hands to the reference's fetch_image."""
    import base64
    import numpy as np
    g = np.random.default_rng(5)
    rgb = Image.fromarray(g.integers(0, 256, (550, 583, 3), dtype=np.uint8))
    rgba = Image.fromarray(g.integers(0, 256, (90, 120, 4), dtype=np.uint8), "RGBA")
    gray = Image.fromarray(g.integers(0, 256, (64, 200), dtype=np.uint8), "L")
    png = os.path.join(tmpdir, "page.png")
    rgb.save(png)
    with open(png, "rb") as f:
        data_url = "data:image/png;base64," + base64.b64encode(f.read()).decode()
    return {"rgb": rgb, "rgba": rgba, "gray": gray, "path": png, "file_url": "file://" + png, "data_url": data_url}


def test_fetch_image_equals_the_reference_function(tmp_path):
    """Loaders (PIL / path / file:// / data: URL), RGB conversion (alpha on white) and the three resize rules: same size,
    mode and pixel bytes as the reference's fetch_image on every case."""
    import zlib
    import pytest
    from dots_ocr_b200.utils.image_utils import fetch_image
    inputs = _fetch_inputs(str(tmp_path))
    assert len(G["fetch"]) == 36
    for case in G["fetch"]:
        r = fetch_image(inputs[case["input"]], **case["kw"])
        assert (r.mode, list(r.size)) == (case["mode"], case["size"]), case
        assert zlib.crc32(r.tobytes()) == case["crc32"], case
    with pytest.raises(ValueError):
        fetch_image(12345)
    with pytest.raises(ValueError):
        fetch_image("data:image/png;hex,00")
