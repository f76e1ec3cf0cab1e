#!/usr/bin/env python
"""Phase timeline of the decode chain kernel (needs a library built with -DCH_TIMING, e.g. tools/build_variant.sh chtime -DCH_TIMING)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import ops
from dots_ocr_b200.engine import _interleave_gate_up
DEV = "cuda:0"
B, H, I, QKV = 64, 1536, 8960, 2048
so, sd, sq = 12, 12, 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=DEV) * sc).to(torch.bfloat16)
attn = r(B, H); w_o, w_down, w_qkv = r(H, H, sc=0.03), r(H, I, sc=0.02), r(QKV, H, sc=0.03)
w_gu = _interleave_gate_up(r(I, H, sc=0.03), r(I, H, sc=0.03))
ln = r(H)
resid, normed, act = r(B, H), torch.empty((B, H), device=DEV, dtype=torch.bfloat16), torch.empty((B, I), device=DEV, dtype=torch.bfloat16)
part = torch.zeros(max(so * H, sd * H, sq * QKV) * B, device=DEV, dtype=torch.float32)
ctr = torch.zeros(16 + 64 + 16, device=DEV, dtype=torch.int32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
names = ["start", "P0 epi done", "P0 arrived", "P1 fin done", "P1 arrived", "P2 epi done", "P2 arrived", "P3 epi done", "P3 arrived",
         "P4 fin done", "P4 arrived", "P5 epi done", "P2 mma first", "P2 mma kb8", "P2 mma last", "P2 tmem_full"]
for it in range(4):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.decode_chain(attn, w_o, w_gu, w_down, w_qkv, part, resid, normed, act, ln, ln, ctr, so, sd, sq, 1e-6)
    e1.record()
    torch.cuda.synchronize()
    st = ctr[16:16 + 64].view(torch.int64).cpu().tolist()
    print(f"iter {it}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us")
    for slot, nm in ((0, "cta 0 (has finalize row)"), (16, f"cta {B} (no finalize row)")):
        t0 = st[slot]
        print("  ", nm, " ".join(f"{names[i]}={(st[slot + i] - t0) / 1e3:.1f}" for i in range(16)))
