"""Generate tests/golden/smart_resize.json by EXECUTING the reference's own function
(/root/reference/dots_ocr/utils/image_utils.py:29-63).  Only runs in the build container
(the reference is not shipped to the GPU box); the JSON it writes is committed."""
import json
import os
import random
import sys
import types

sys.path.insert(0, "/root/reference")
sys.modules["fitz"] = types.ModuleType("fitz")          # PyMuPDF is not installed; not needed for this function
from dots_ocr.utils.image_utils import smart_resize      # noqa: E402

rng = random.Random(20260922)
cases = [(1024, 1024), (1960, 1960), (2250, 1700), (583, 550), (946, 1024), (28, 28), (27, 5000), (14, 14), (1, 150),
         (3360, 3360), (3361, 3361), (5000, 5000), (10000, 60), (56, 11200), (200, 1), (100, 20001)]
for _ in range(400):
    cases.append((rng.randint(1, 6000), rng.randint(1, 6000)))
out = []
for h, w in cases:
    for kw in ({}, {"min_pixels": 3136, "max_pixels": 1003520}, {"min_pixels": 200704, "max_pixels": 11289600}):
        try:
            r = list(smart_resize(h, w, **kw))
        except ValueError:
            r = "ValueError"
        out.append({"h": h, "w": w, "kw": kw, "out": r})
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smart_resize.json")
json.dump(out, open(path, "w"))
print(len(out), "cases ->", path)
