"""Golden vectors of the tiny-config oracle (CPU fp32, seeded synthetic weights and inputs).
Committed as tests/golden/oracle_tiny.npz; regenerate with `python tests/golden/make_oracle_golden.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dots_ocr_b200 import config, weights   # noqa: E402
from oracle.model import DotsOracle          # noqa: E402

GRIDS = [(1, 8, 8), (1, 6, 10)]
N_NEW = 12


def golden_inputs(cfg, seed=99):
    g = torch.Generator().manual_seed(seed)
    pvs, rows = [], []
    for (t, h, w) in GRIDS:
        S = t * h * w
        pvs.append(torch.randn(S, cfg.vision.patch_dim, generator=g))
        rows.append(torch.cat([torch.randint(0, 2000, (4,), generator=g), torch.full((S // 4,), cfg.image_token_id),
                               torch.randint(0, 2000, (6,), generator=g)]))
    T = max(r.numel() for r in rows)
    ids = torch.zeros((len(rows), T), dtype=torch.long)
    mask = torch.zeros_like(ids)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    return torch.cat(pvs), torch.tensor(GRIDS), ids, mask


def main():
    torch.set_num_threads(1)
    cfg = config.tiny()
    pv, grid, ids, mask = golden_inputs(cfg)
    out = {}
    for fl in ("peaked", "random"):
        ck = weights.make_synthetic_checkpoint(cfg, 0, fl)
        o = DotsOracle(cfg, ck, torch.float32, "cpu")
        seq = o.generate(ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N_NEW)
        img = o.vision.forward(pv, grid)
        out[f"{fl}_sequences"] = seq.numpy()
        out[f"{fl}_image_embeds"] = img.numpy().astype(np.float32)
        if fl == "random":
            # row 0 is un-padded?  use per-row teacher forcing on the unpadded rows
            for b in range(ids.shape[0]):
                keep = mask[b].bool()
                lg = o.teacher_forced_logits(ids[b][keep].unsqueeze(0), seq[b, ids.shape[1]:].unsqueeze(0),
                                             pv[: 64] if b == 0 else pv[64:], grid[b:b + 1])
                top = lg[0].topk(8, -1)
                out[f"random_top8_val_{b}"] = top.values.numpy()
                out[f"random_top8_idx_{b}"] = top.indices.numpy()
                out[f"random_logit_std_{b}"] = np.array(float(lg.std()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_tiny.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
