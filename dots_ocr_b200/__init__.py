"""dots_ocr_b200: B200-native engine for the dots.ocr page-parsing hot path.

Public surface mirrors the reference package (``dots_ocr/__init__.py:1``):
``from dots_ocr_b200 import DotsOCRParser``; the engine itself is ``dots_ocr_b200.engine.Engine``.
"""
from .parser import DotsOCRParser  # noqa: F401

__all__ = ["DotsOCRParser"]
