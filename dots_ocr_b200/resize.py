"""Bicubic + antialias resize of uint8 pages on the GPU, bit-identical to the resize the stock image processor runs on the CPU.

The reference resizes every page to a multiple of 28 inside the pixel budget before the model sees it
(``dots_ocr/utils/image_utils.py:116-138`` fetch_image -> PIL; the HF fast processor repeats it with
``torchvision.transforms.v2.functional.resize(uint8, BICUBIC, antialias=True)``,
``transformers/models/qwen2_vl/image_processing_qwen2_vl.py:148-232``).  Both run the same algorithm (Pillow's
ImagingResample, ported to ATen for uint8 tensors): two separable passes -- horizontal first, then vertical, with a uint8
intermediate image -- whose filter taps are computed in double precision, normalised, and quantised to int16 fixed point;
each output sample is ``clip((2^(p-1) + sum_j in[xmin + j] * w[j]) >> p, 0, 255)`` in int32.

This module builds the per-axis tap tables on the host (a few KB, cached per (in, out) size pair); ``dots_resize_bicubic_u8``
applies them on the device in integer arithmetic, so the result equals the CPU resize bit for bit
(tests/test_resize.py checks the restatement against torchvision here, tests/test_ops_gpu.py the kernel on the GPU).
SURVEY.md section 8f N1.
"""
from __future__ import annotations

import functools
import math
from typing import Tuple

import numpy as np


def _bicubic(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0, np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


@functools.lru_cache(maxsize=256)
def axis_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """Taps of one axis: (xmin [out] int32, xsize [out] int32, weights [out, ksize] int16, precision bits)."""
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    i = np.arange(out_size, dtype=np.float64)
    center = scale * (i + 0.5)
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)             # C cast: the operands are positive here
    hi = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = hi - lo
    j = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((j + lo[:, None] - center[:, None] + 0.5) * invscale)
    w = np.where(j < n[:, None], w, 0.0)
    # normalise with the running sum in tap order, as the C loop accumulates it
    tot = np.zeros(out_size, dtype=np.float64)
    for k in range(ksize):
        tot = tot + w[:, k]
    w = np.where(tot[:, None] != 0.0, w / np.where(tot == 0.0, 1.0, tot)[:, None], w)
    wmax = float(w.max())
    prec = 0
    while prec < 22:
        if int(0.5 + wmax * (1 << (prec + 1))) >= (1 << 15):
            break
        prec += 1
    wi = np.trunc(np.where(w < 0, -0.5 + w * (1 << prec), 0.5 + w * (1 << prec))).astype(np.int16)
    return lo.astype(np.int32), n.astype(np.int32), np.ascontiguousarray(wi), prec


def resize_u8_reference(img: np.ndarray, rh: int, rw: int) -> np.ndarray:
    """CPU restatement of the two-pass integer resize (test infrastructure for the tables; [H, W, C] uint8 in and out)."""
    def one_axis(x: np.ndarray, axis: int, out_size: int) -> np.ndarray:
        lo, n, wi, prec = axis_tables(x.shape[axis], out_size)
        x = np.moveaxis(x, axis, 0).astype(np.int32)
        out = np.empty((out_size,) + x.shape[1:], np.uint8)
        for i in range(out_size):
            acc = np.full(x.shape[1:], 1 << (prec - 1), np.int32)
            for k in range(int(n[i])):
                acc = acc + x[lo[i] + k] * int(wi[i, k])
            out[i] = np.clip(acc >> prec, 0, 255)
        return np.moveaxis(out, 0, axis)
    H, W = img.shape[:2]
    x = img
    if W != rw:
        x = one_axis(x, 1, rw)
    if H != rh:
        x = one_axis(x, 0, rh)
    return np.ascontiguousarray(x)
