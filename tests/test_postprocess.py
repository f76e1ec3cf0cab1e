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
