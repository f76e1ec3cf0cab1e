"""Layout cells -> Markdown.  Same call surface as ``dots_ocr/utils/format_transformer.py``
(``has_latex_markdown`` :10-38, ``clean_latex_preamble`` :41-65, ``get_formula_in_markdown`` :68-119, ``clean_text`` :122-142,
``layoutjson2md`` :145-180, ``fix_streamlit_formulas`` :183-206); behaviour pinned by ``tests/golden/postprocess.json``."""
from __future__ import annotations

import base64
import re
from io import BytesIO

_LATEX_MARKS = re.compile(
    r"\$\$.*?\$\$"                       # $$ ... $$
    r"|\$[^$\n]+?\$"                     # $ ... $
    r"|\\begin\{.*?\}.*?\\end\{.*?\}"    # environments
    r"|\\[a-zA-Z]+\{.*?\}"               # \command{...}
    r"|\\[a-zA-Z]+"                      # \command
    r"|\\\[.*?\\\]"                      # \[ ... \]
    r"|\\\(.*?\\\)",                     # \( ... \)
    re.DOTALL)
_PREAMBLE = re.compile(
    r"\\documentclass\{[^}]+\}|\\usepackage\{[^}]+\}|\\usepackage\[[^\]]*\]\{[^}]+\}|\\begin\{document\}|\\end\{document\}",
    re.IGNORECASE)
_INLINE_DOLLAR = re.compile(r"\$([^$]+)\$")
_HAS_BRACKET_MATH = re.compile(r".*\\\[.*\\\].*")
_BLOCK = re.compile(r"\$\$(.*?)\$\$", re.DOTALL)


def PILimage_to_base64(image, format="PNG") -> str:
    buf = BytesIO()
    image.save(buf, format=format)
    return f"data:image/{format.lower()};base64,{base64.b64encode(buf.getvalue()).decode('utf-8')}"


def has_latex_markdown(text) -> bool:
    return isinstance(text, str) and _LATEX_MARKS.search(text) is not None


def clean_latex_preamble(latex_text: str) -> str:
    return _PREAMBLE.sub("", latex_text)


def get_formula_in_markdown(text: str) -> str:
    """A formula cell's text as a Markdown display block (already-delimited input is normalised, inline math is kept)."""
    text = text.strip()
    if text.startswith("$$") and text.endswith("$$"):
        inner = text[2:-2].strip()
        return text if "$" in inner else f"$$\n{inner}\n$$"
    if text.startswith("\\[") and text.endswith("\\]"):
        return f"$$\n{text[2:-2].strip()}\n$$"
    if _HAS_BRACKET_MATH.findall(text):
        return text
    if _INLINE_DOLLAR.findall(text):
        return text
    if not has_latex_markdown(text):
        return text
    if "usepackage" in text:
        text = clean_latex_preamble(text)
    if text[0] == "`" and text[-1] == "`":
        text = text[1:-1]
    return f"$$\n{text}\n$$"


def clean_text(text: str) -> str:
    if not text:
        return ""
    text = text.strip()
    if text[:2] == "`$" and text[-2:] == "$`":
        text = text[1:-1]
    return text


def layoutjson2md(image, cells: list, text_key: str = "text", no_page_hf: bool = False) -> str:
    """Formulas are LaTeX, tables HTML, text Markdown; pictures are cropped from `image` and inlined as base64 PNG."""
    parts = []
    for cell in cells:
        x1, y1, x2, y2 = (int(v) for v in cell["bbox"])
        category = cell["category"]
        if no_page_hf and category in ("Page-header", "Page-footer"):
            continue
        if category == "Picture":
            parts.append(f"![]({PILimage_to_base64(image.crop((x1, y1, x2, y2)))})")
        elif category == "Formula":
            parts.append(get_formula_in_markdown(cell.get(text_key, "")))
        else:
            parts.append(clean_text(cell.get(text_key, "")))
    return "\n\n".join(parts)


def fix_streamlit_formulas(md: str) -> str:
    def block(m):
        body = m.group(1)
        if body.startswith("\n"):
            body = body[1:]
        if body.endswith("\n"):
            body = body[:-1]
        return f"$$\n{body}\n$$"
    return _BLOCK.sub(block, md)
