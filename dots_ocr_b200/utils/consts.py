"""Pixel-budget constants of the dots.ocr pre-processor.

Mirrors ``dots_ocr/utils/consts.py:1-5`` of the reference (values only).
"""
MIN_PIXELS = 3136          # 4 merge blocks of 28x28
MAX_PIXELS = 11289600      # 3360 x 3360
IMAGE_FACTOR = 28          # patch 14 x spatial merge 2

image_extensions = {".jpg", ".jpeg", ".png"}
