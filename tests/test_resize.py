"""The tap tables of the GPU page resize (dots_ocr_b200/resize.py) reproduce torchvision's uint8 bicubic + antialias resize -- the
resize of the stock image processor (image_processing_qwen2_vl.py:148-232) -- bit for bit, up- and down-scaling."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("H,W,rh,rw", [(64, 80, 56, 84), (100, 100, 112, 112), (50, 70, 28, 56), (300, 200, 140, 84), (33, 47, 56, 56),
                                       (257, 311, 252, 308), (120, 90, 336, 252), (1024, 1024, 1036, 1036), (90, 64, 90, 56), (61, 70, 56, 70)])
def test_restated_resize_equals_torchvision(H, W, rh, rw):
    import torchvision.transforms.v2.functional as tvF
    from torchvision.transforms import InterpolationMode
    from dots_ocr_b200.resize import resize_u8_reference
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if (H, W) == (100, 100):
        img[:] = np.where(rng.random((H, W, 1)) < 0.5, 0, 255)           # saturated edges: exercises the clamp
    ref = tvF.resize(torch.from_numpy(img).permute(2, 0, 1).contiguous(), [rh, rw], interpolation=InterpolationMode.BICUBIC,
                     antialias=True).permute(1, 2, 0).numpy()
    assert np.array_equal(resize_u8_reference(img, rh, rw), ref)


def test_tables_shape_and_bounds():
    from dots_ocr_b200.resize import axis_tables
    for a, b in [(1024, 1036), (2000, 1008), (28, 3360), (3000, 56)]:
        lo, n, w, prec = axis_tables(a, b)
        assert lo.shape == n.shape == (b,) and w.shape[0] == b and w.dtype == np.int16
        assert (lo >= 0).all() and (lo + n <= a).all() and (n >= 1).all() and (n <= w.shape[1]).all()
        assert 0 < prec <= 22 and (np.abs(w.astype(np.int64).sum(1) - (1 << prec)) <= w.shape[1]).all()
