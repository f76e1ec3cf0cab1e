"""Generate tests/golden/postprocess.json by EXECUTING the reference's own post-processing functions
(/root/reference/dots_ocr/utils/layout_utils.py:115-228, format_transformer.py:10-206).  Only runs in the build container
(the reference is not shipped to the GPU box); the JSON it writes is committed."""
import json
import os
import random
import sys
import types

sys.path.insert(0, "/root/reference")
sys.modules["fitz"] = types.ModuleType("fitz")          # PyMuPDF is not installed; these functions do not use it
from PIL import Image                                     # noqa: E402
from dots_ocr.utils import layout_utils as L              # noqa: E402
from dots_ocr.utils import format_transformer as F        # noqa: E402

rng = random.Random(7)
out = {"cells": [], "bboxes": [], "formula": [], "clean_text": [], "has_latex": [], "preamble": [], "md": [], "fix": [], "output": []}

sizes = [(1700, 2250), (1024, 1024), (583, 550), (3000, 200), (4000, 4000), (100, 60)]
for (ow, oh) in sizes:
    img = Image.new("RGB", (ow, oh), (255, 255, 255))
    for (iw, ih) in [(ow, oh), (ow // 2 + 3, oh // 2 + 1), (1036, 1036)]:
        for kw in ({}, {"min_pixels": 3136, "max_pixels": 1003520}, {"min_pixels": None, "max_pixels": None}):
            cells = [{"bbox": [rng.randint(0, iw), rng.randint(0, ih), rng.randint(0, iw), rng.randint(0, ih)], "category": "Text", "text": f"t{i}"}
                     for i in range(5)]
            cells.append({"bbox": [1.5, "2", 3.9, 4.2], "category": "Title"})
            out["cells"].append({"origin": [ow, oh], "input": [iw, ih], "kw": kw, "cells": cells,
                                 "out": L.post_process_cells(img, cells, iw, ih, **kw), "legal": L.is_legal_bbox(cells)})
            bbs = [[rng.randint(0, ow), rng.randint(0, oh), rng.randint(0, ow), rng.randint(0, oh)] for _ in range(4)]
            out["bboxes"].append({"origin": [ow, oh], "input": [iw, ih], "kw": kw, "bboxes": bbs,
                                  "out": L.pre_process_bboxes(img, bbs, iw, ih, **kw)})

formulas = ["$$ a+b $$", "$$a$b$$", "\\[ x^2 \\]", "see \\[ x \\] here", "$x$ and $y$", "plain text", "\\frac{a}{b}", "  \\alpha  ",
            "\\usepackage{amsmath}\\begin{document}E=mc^2\\end{document}", "`\\sum_i x_i`", "$$\n\\int f\n$$", "x = 1", "\\(a\\)",
            "\\begin{aligned}a&=b\\end{aligned}", "$$$$", "`a`"]
for t in formulas:
    out["formula"].append({"in": t, "out": F.get_formula_in_markdown(t)})
    out["has_latex"].append({"in": t, "out": F.has_latex_markdown(t)})
    out["preamble"].append({"in": t, "out": F.clean_latex_preamble(t)})
out["has_latex"].append({"in": None, "out": F.has_latex_markdown(None)})
for t in ["", None, "  hi  ", "`$x$`", "`$x$", "a\n b", "`$`"]:
    out["clean_text"].append({"in": t, "out": F.clean_text(t)})
for t in ["$$a$$", "x $$\na\n$$ y $$b\n$$", "no math", "$$\n\nq\n\n$$"]:
    out["fix"].append({"in": t, "out": F.fix_streamlit_formulas(t)})

img = Image.new("RGB", (64, 48), (10, 200, 30))
cells = [{"bbox": [0, 0, 10, 10], "category": "Page-header", "text": " head "}, {"bbox": [1, 2, 30, 20], "category": "Formula", "text": "\\frac{1}{2}"},
         {"bbox": [5, 5, 20, 25], "category": "Picture"}, {"bbox": [0, 30, 64, 48], "category": "Table", "text": "<table><tr><td>1</td></tr></table>"},
         {"bbox": [0, 40, 64, 48], "category": "Page-footer", "text": "3"}, {"bbox": [0, 0, 1, 1], "category": "Text", "text": "`$z$`"},
         {"bbox": [0, 0, 1, 1], "category": "Text"}]
for no_hf in (False, True):
    out["md"].append({"size": [64, 48], "color": [10, 200, 30], "cells": cells, "no_page_hf": no_hf, "out": F.layoutjson2md(img, cells, no_page_hf=no_hf)})

page, seen = Image.new("RGB", (1700, 2250)), Image.new("RGB", (1708, 2240))
good = json.dumps([{"bbox": [10, 20, 300, 400], "category": "Text", "text": "hello"}])
for mode, resp in [("prompt_layout_all_en", good), ("prompt_ocr", "raw text"), ("prompt_layout_only_en", good), ("prompt_table_html", "<table/>")]:
    r = L.post_process_output(resp, mode, page, seen)
    out["output"].append({"mode": mode, "response": resp, "out": list(r) if isinstance(r, tuple) else r, "tuple": isinstance(r, tuple)})

# ---- OutputCleaner (dots_ocr/utils/output_cleaner.py:32-435) and the failure path of post_process_output -------------------
import contextlib
import io
from dots_ocr.utils.output_cleaner import OutputCleaner   # noqa: E402

def cell(i, cat="Text", text=None, bbox=None):
    d = {"bbox": bbox or [i, i + 1, i + 10, i + 20], "category": cat}
    if text is not None:
        d["text"] = text
    return d

full = json.dumps([cell(i, text=f"line {i}") for i in range(6)])
cases = [
    full,
    full[:-1],                                            # array never closed
    full[: len(full) - 25],                               # cut inside the last cell
    full[: full.find('"text"') + 12],                     # cut inside the first cell's text
    '[{"bbox": [1, 2, 3, 4], "category": "Title", "text": "only one and it is cut',
    '[{"bbox": [1, 2, 3], "category": "Title", "text": "three coords" ',
    '[{"bbox": [1, 2, x, 4], "category": "Title", "text": "bad int',
    '{"bbox": [1, 2, 3, 4], "category": "Text", "text": "a"}{"bbox": [5, 6, 7, 8], "category": "Text", "text": "b"}',
    '[{"bbox": [1, 2, 3, 4], "category": "Text", "text": "a"} {"bbox": [5, 6, 7, 8], "category": "Text", "text": "b"}]',
    json.dumps([cell(1, text="same")] * 4 + [cell(9, text="other")]),
    json.dumps([cell(i, text="spam") for i in range(7)] + [cell(50, text="tail")]),
    json.dumps([cell(i, text="spam") for i in range(4)]),
    json.dumps([cell(3, text="a"), cell(3, text="b"), cell(4, text="c")]),
    "not json at all",
    "",
    "[]",
    '{"bbox": [1, 2, 3, 4], "category": "Text"}',
    '[{"bbox": [1,2,3,4], "category": "Table", "text": "<table><tr><td>{x}</td></tr></table>"}, {"bbox": [5,6,7,8], "category": "Text", "text": "y"}]',
    '[{"bbox": [1,2,3,4], "category": "Formula", "text": "a_{i}"}, {"bbox": [5,6,7,8], "category": "Text", "text": "z"',
    "[" + ", ".join(json.dumps(cell(i, text="w" * 400)) for i in range(140)),            # > 50 000 chars, not closed
    "[" + ", ".join(json.dumps(cell(i, text="w" * 400)) for i in range(140)) + "]",      # > 50 000 chars, closed: still loses its last cell
    '[{"category": "Text", "text": "no bbox"}, {"bbox": [1,2,3,4], "category": "Text", "text": "ok"}]',
    "42", "null", '"a string"',
]
list_cases = [
    [cell(1, text="a"), {"bbox": [1, 2, 3], "category": "Text", "text": "three"}, {"bbox": [1, 2, 3]}, {"bbox": "oops", "category": "X"},
     {"category": "Picture"}, {"text": "orphan"}, "junk", 7, cell(1, text="dup bbox")],
    [],
    [cell(i, text="spam") for i in range(6)],
    [{"bbox": [[1, 2], 3, 4, 5], "category": "Text", "text": "nested"}, {"bbox": [[1, 2], 3, 4, 5], "category": "Text", "text": "nested2"}],
    [cell(1)],
]
out["cleaner"] = []
for c in cases + list_cases:
    with contextlib.redirect_stdout(io.StringIO()):
        r = OutputCleaner().clean_model_output(c)
    out["cleaner"].append({"in": c, "out": r})

out["output_fail"] = []
for resp in [cases[2], cases[3], cases[4], "not json", "[]", json.dumps([{"category": "Text", "text": "no bbox here"}]), json.dumps({"bbox": [1, 2, 3, 4]}),
             cases[10], json.dumps([cell(1, text="a"), {"bbox": [1, 2, 3], "category": "Text", "text": "three"}])]:
    with contextlib.redirect_stdout(io.StringIO()):
        r = L.post_process_output(resp, "prompt_layout_all_en", page, seen)
    out["output_fail"].append({"response": resp, "out": list(r)})

# ---- fetch_image (dots_ocr/utils/image_utils.py:84-138): loaders, RGB conversion, the three resize rules ---------------------
import base64          # noqa: E402
import tempfile        # noqa: E402
import zlib            # noqa: E402
import numpy as np     # noqa: E402
from dots_ocr.utils import image_utils as IU   # noqa: E402


def fetch_inputs(tmpdir):
    """name -> what is handed to fetch_image (shared with the test, which rebuilds the same inputs)."""
    g = np.random.default_rng(5)
    rgb = Image.fromarray(g.integers(0, 256, (550, 583, 3), dtype=np.uint8))
    rgba = Image.fromarray(g.integers(0, 256, (90, 120, 4), dtype=np.uint8), "RGBA")
    gray = Image.fromarray(g.integers(0, 256, (64, 200), dtype=np.uint8), "L")
    png = os.path.join(tmpdir, "page.png")
    rgb.save(png)
    with open(png, "rb") as f:
        data_url = "data:image/png;base64," + base64.b64encode(f.read()).decode()
    return {"rgb": rgb, "rgba": rgba, "gray": gray, "path": png, "file_url": "file://" + png, "data_url": data_url}


FETCH_KW = [{}, {"min_pixels": 3136, "max_pixels": 200704}, {"max_pixels": 100352}, {"min_pixels": 1003520},
            {"resized_height": 300, "resized_width": 500}, {"resized_height": 30, "resized_width": 45, "min_pixels": 3136}]
out["fetch"] = []
with tempfile.TemporaryDirectory() as td:
    for name, src in fetch_inputs(td).items():
        for kw in FETCH_KW:
            r = IU.fetch_image(src, **kw)
            out["fetch"].append({"input": name, "kw": kw, "mode": r.mode, "size": list(r.size), "crc32": zlib.crc32(r.tobytes())})

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "postprocess.json")
json.dump(out, open(path, "w"))
print({k: len(v) for k, v in out.items()}, "->", path)
