"""PDF pages -> PIL images (the host step in front of the hot path for ``.pdf`` inputs).

Behaviour of the reference's ``load_images_from_pdf`` / ``fitz_doc_to_image`` / ``get_image_by_fitz_doc``
(``dots_ocr/utils/doc_utils.py:19-60``, ``dots_ocr/utils/image_utils.py:169-199``): every page in
[start_page_id, end_page_id] is rendered at ``dpi`` (scale dpi/72, no alpha); a page whose render would exceed 4500 px on
either side is rendered at PyMuPDF's native 72 dpi instead.

PyMuPDF is the reference's dependency for this step and is NOT in this image: it is imported when a PDF is actually
opened, and its absence raises ``RasteriserUnavailable`` (an ImportError) naming the workaround -- rasterise elsewhere and
hand the page images to ``DotsOCRParser.parse_pages``.  The tests drive this module with a stand-in ``fitz``.
"""
from __future__ import annotations

from typing import List, Optional

MAX_RENDER_SIDE = 4500      # doc_utils.py:34


class RasteriserUnavailable(ImportError):
    pass


def _fitz():
    try:
        import fitz          # PyMuPDF
    except ImportError:
        try:
            import pymupdf as fitz
        except ImportError:
            raise RasteriserUnavailable(
                "PDF input needs PyMuPDF (`fitz`), which is not installed; rasterise the pages yourself and call "
                "DotsOCRParser.parse_pages(images, ...)") from None
    return fitz


def render_page(page, dpi: int = 200):
    """One PyMuPDF page -> RGB PIL image at ``dpi`` (72 dpi if the ``dpi`` render is wider or taller than 4500 px)."""
    from PIL import Image
    fitz = _fitz()
    scale = dpi / 72.0
    pix = page.get_pixmap(matrix=fitz.Matrix(scale, scale), alpha=False)
    if max(pix.width, pix.height) > MAX_RENDER_SIDE:
        pix = page.get_pixmap(matrix=fitz.Matrix(1.0, 1.0), alpha=False)
    return Image.frombytes("RGB", (pix.width, pix.height), pix.samples)


def load_images_from_pdf(pdf_file: str, dpi: int = 200, start_page_id: int = 0, end_page_id: Optional[int] = None) -> List:
    fitz = _fitz()
    with fitz.open(pdf_file) as doc:
        last = doc.page_count - 1
        if end_page_id is None or end_page_id < 0 or end_page_id > last:
            end_page_id = last
        return [render_page(doc[i], dpi) for i in range(max(0, start_page_id), end_page_id + 1)]


def get_image_by_fitz_doc(image, target_dpi: int = 200):
    """Re-render an image through a one-page PDF at ``target_dpi`` (the reference's ``fitz_preprocess`` option for
    low-resolution scans)."""
    import io
    from PIL import Image
    fitz = _fitz()
    if isinstance(image, Image.Image):
        buf = io.BytesIO()
        image.save(buf, format="PNG")
        data = buf.getvalue()
    else:
        with open(image, "rb") as f:
            data = f.read()
    with fitz.open(stream=data) as src:
        pdf_bytes = src.convert_to_pdf()
    with fitz.open("pdf", pdf_bytes) as doc:
        return render_page(doc[0], target_dpi)
