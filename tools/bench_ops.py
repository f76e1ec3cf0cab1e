"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between timed launches).
Prints one JSON line per kernel/shape; used to steer optimisation, not as the headline bench."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dots_ocr_b200 import ops  # noqa: E402
from dots_ocr_b200.engine import _interleave_gate_up  # noqa: E402

DEV = "cuda:0"
flush = None


def timeit(fn, iters=5, warmup=2):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(torch.bfloat16)


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    only = sys.argv[1:] or ["gemm", "skinny", "attn", "decode", "norm"]
    if "gemm" in only:
        for (M, N, K, epi) in [(5476 * 8, 4608, 1536, "store"), (5476 * 8, 1536, 1536, "res"), (5476 * 8, 8448, 1536, "swiglu"),
                               (5476 * 8, 1536, 4224, "res"), (1625 * 16, 2048, 1536, "bias"), (1625 * 16, 17920, 1536, "swiglu"),
                               (1625 * 16, 1536, 8960, "res"), (1369 * 8, 6144, 6144, "gelu"), (8192, 8192, 8192, "store")]:
            a, w = rnd(M, K), rnd(N, K, scale=0.03)
            if epi == "store":
                f = lambda: ops.gemm(a, w)
            elif epi == "bias":
                b = rnd(N); f = lambda: ops.gemm(a, w, epilogue=ops.EPI_BIAS, bias=b)
            elif epi == "gelu":
                b = rnd(N); f = lambda: ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=b)
            elif epi == "res":
                r = rnd(M, N); f = lambda: ops.gemm(a, w, out=r, epilogue=ops.EPI_RESIDUAL, residual=r)
            else:
                f = lambda: ops.gemm(a, w, epilogue=ops.EPI_SWIGLU)
            ops.set_gemm_pair(False)
            ms = timeit(f)
            ops.set_gemm_pair(True)
            ms2 = timeit(f)
            tf = 2.0 * M * N * K / ms / 1e9
            ref = timeit(lambda: torch.matmul(a, w.t()))
            print(json.dumps(dict(k="gemm", M=M, N=N, K=K, epi=epi, ms=round(ms, 4), tflops=round(tf, 1), pair_ms=round(ms2, 4),
                                  pair_tflops=round(2.0 * M * N * K / ms2 / 1e9, 1),
                                  cublas_ms=round(ref, 4), cublas_tflops=round(2.0 * M * N * K / ref / 1e9, 1))), flush=True)
            del a, w
    if "decode_gemm" in only:
        # decode GEMMs over pre-tiled operands (bulk copies) vs the tensor-map kernels, weights >> L2 are not possible for one layer, so
        # each timing loops over 8 distinct weight copies (the step streams a different layer's weights every time)
        for (B, N, K) in [(64, 2048, 1536), (64, 1536, 1536), (64, 17920, 1536), (64, 1536, 8960), (64, 151936, 1536)]:
            copies = 1 if N > 100000 else 8
            ws = [rnd(N, K, scale=0.03) for _ in range(copies)]
            wts = [ops.tile_weight(w) for w in ws]
            x = rnd(B, K)
            xt = ops.tile_rows(x, 64).view(-1)
            s_ = 1 if N > 10000 else ops.pick_splits(-(-N // 128), -(-K // 64))
            part = torch.empty((s_, B, N), device=DEV, dtype=torch.float32)
            it = [0]

            def old():
                it[0] += 1
                ops.gemm_skinny(x, ws[it[0] % copies], s_, partial=part)

            def new():
                it[0] += 1
                ops.decode_gemm_partial(xt, wts[it[0] % copies], part, B, N, K, s_)
            for name, fn in (("tensor-map", old), ("bulk-tiled", new)):
                ms = timeit(fn, iters=16)
                print(json.dumps(dict(k="decode_gemm", path=name, B=B, N=N, K=K, splits=s_, ms=round(ms, 4), weight_GBs=round(N * K * 2 / ms / 1e6, 1))), flush=True)
            del ws, wts
    if "skinny" in only:
        for (B, N, K) in [(64, 2048, 1536), (64, 1536, 1536), (64, 17920, 1536), (64, 1536, 8960), (64, 151936, 1536), (1, 17920, 1536)]:
            x, w = rnd(B, K), rnd(N, K, scale=0.03)
            s = 1 if N > 10000 else ops.pick_splits(-(-N // 128), -(-K // 64))
            part = torch.empty((s, B, N), device=DEV, dtype=torch.float32)
            ms = timeit(lambda: ops.gemm_skinny(x, w, s, partial=part), iters=9)
            gbs = (N * K * 2) / ms / 1e6
            print(json.dumps(dict(k="skinny", B=B, N=N, K=K, splits=s, ms=round(ms, 4), weight_GBs=round(gbs, 1))), flush=True)
    if "attn" in only:
        for (L, nseq, hq, hkv, causal) in [(5476, 4, 12, 12, False), (19600, 1, 12, 12, False), (1625, 16, 12, 2, True)]:
            T = L * nseq
            qkv = rnd(T, (hq + 2 * hkv) * 128)
            out = torch.empty((T, hq * 128), device=DEV, dtype=torch.bfloat16)
            cu = torch.arange(0, nseq + 1, device=DEV, dtype=torch.int32) * L
            q, k, v = qkv[:, : hq * 128], qkv[:, hq * 128:(hq + hkv) * 128], qkv[:, (hq + hkv) * 128:]
            fl = 4.0 * L * L * 128 * hq * nseq * (0.5 if causal else 1.0)
            for impl in (("tc", "pair") if ops.has_experiments() else ("tc",)):
                try:
                    ms = timeit(lambda: ops.attn_varlen(q, k, v, out, cu, L, hq, hkv, causal, 128 ** -0.5, impl=impl))
                except Exception as e:
                    print(json.dumps(dict(k="attn", impl=impl, error=repr(e)[:200])), flush=True)
                    continue
                print(json.dumps(dict(k="attn", impl=impl, L=L, nseq=nseq, hq=hq, hkv=hkv, causal=causal, ms=round(ms, 4),
                                      tflops=round(fl / ms / 1e9, 1))), flush=True)
            # the bar: the kernel the reference's HF path really runs (parser.py:68-74 attn_implementation="flash_attention_2";
            # [V] dots_ocr.py:304-310 flash_attn_varlen_func) -- the in-image flash-attn 2.8 build (library code, mma.sync on sm_100)
            try:
                from flash_attn import flash_attn_varlen_func
                q3, k3, v3 = q.reshape(T, hq, 128), k.reshape(T, hkv, 128), v.reshape(T, hkv, 128)
                ms = timeit(lambda: flash_attn_varlen_func(q3, k3, v3, cu, cu, L, L, softmax_scale=128 ** -0.5, causal=causal))
                print(json.dumps(dict(k="attn", impl="flash_attn_2.8_varlen (reference's kernel)", L=L, nseq=nseq, hq=hq, hkv=hkv, causal=causal,
                                      ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))), flush=True)
            except Exception as e:      # noqa: BLE001
                print(json.dumps(dict(k="attn", impl="flash_attn", error=repr(e)[:200])), flush=True)
    if "decode" in only:
        for (B, ctx, splits) in [(64, 1881, 3), (64, 1881, 1), (64, 1881, 6), (1, 1881, 16), (32, 6200, 6)]:
            hq, hkv = 12, 2
            ctx_max = (ctx + 64 + 63) // 64 * 64
            q = rnd(B, hq * 128)
            kc, vc = rnd(B, hkv, ctx_max, 128), rnd(B, hkv, ctx_max, 128)
            cl = torch.full((B,), ctx, device=DEV, dtype=torch.int32)
            out = torch.empty_like(q)
            po = torch.empty((B, hq, splits, 128), device=DEV, dtype=torch.float32)
            pm = torch.empty((B, hq, splits, 2), device=DEV, dtype=torch.float32)
            ms = timeit(lambda: ops.attn_decode(q, kc, vc, cl, out, hq, hkv, ctx_max, splits, 128 ** -0.5, po, pm), iters=9)
            gbs = (2.0 * B * hkv * ctx * 256) / ms / 1e6
            print(json.dumps(dict(k="attn_decode", B=B, ctx=ctx, splits=splits, ms=round(ms, 4), kv_GBs=round(gbs, 1))), flush=True)
    if "norm" in only:
        rows = 5476 * 16
        x, w = rnd(rows, 1536), rnd(1536)
        out = torch.empty_like(x)
        ms = timeit(lambda: ops.rmsnorm(x, w, 1e-5, out=out))
        print(json.dumps(dict(k="rmsnorm", rows=rows, ms=round(ms, 4), GBs=round(2.0 * rows * 1536 * 2 / ms / 1e6, 1))), flush=True)
    print(json.dumps(dict(peaks=peaks)))


if __name__ == "__main__":
    main()
