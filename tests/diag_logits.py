"""Diagnostic (not collected by pytest): per-step logit error of the CUDA engine against the oracle on the tiny config.
Lives under tests/ because it imports the oracle.  Usage: python tests/diag_logits.py [random|peaked]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dots_ocr_b200 import config, weights
from dots_ocr_b200.engine import Engine
from oracle.model import DotsOracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_engine_gpu import _inputs
DEV = "cuda:0"
cfg = config.tiny()
fl = sys.argv[1] if len(sys.argv) > 1 else "random"
ck = weights.make_synthetic_checkpoint(cfg, 0, fl)
eng = Engine(cfg, ck, DEV)
pv, grid, rows = _inputs(cfg, [(1, 8, 8), (1, 8, 8), (1, 8, 8)])
ids = torch.stack(rows)
N = 8
o32 = DotsOracle(cfg, ck, torch.float32, "cpu")
o16 = DotsOracle(cfg, ck, torch.bfloat16, DEV)
ref_ids = o32.generate(ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N)
new = ref_ids[:, ids.shape[1]:]
ref = o32.teacher_forced_logits(ids, new, pv, grid)
ref16 = o16.teacher_forced_logits(ids, new, pv.to(DEV), grid).cpu()
out = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N, forced_ids=new, return_logits=True)
got = out.logits.float().cpu()
sd = ref.std()
print("std", float(sd))
for s in range(N):
    print(s, "eng-vs-fp32", float((got[:, s] - ref[:, s]).abs().max() / sd), "hf16-vs-fp32", float((ref16[:, s] - ref[:, s]).abs().max() / sd),
          "eng-vs-hf16", float((got[:, s] - ref16[:, s]).abs().max() / sd))
# image embeds
img32 = o32.vision.forward(pv, grid)
print("img embeds err", float((out.image_embeds.float().cpu() - img32).abs().max() / img32.abs().max()))
# text-only prompt
ids2 = torch.randint(0, 2000, (2, 12))
r2 = o32.generate(ids2, max_new_tokens=4)
n2 = r2[:, 12:]
l32 = o32.teacher_forced_logits(ids2, n2)
l16 = o16.teacher_forced_logits(ids2, n2).cpu()
g2 = eng.generate(ids2, max_new_tokens=4, forced_ids=n2, return_logits=True).logits.float().cpu()
for s in range(4):
    print("text-only", s, float((g2[:, s] - l32[:, s]).abs().max() / l32.std()), float((l16[:, s] - l32[:, s]).abs().max() / l32.std()))
