"""One worker process per GPU behind the one-page call surface (dots_ocr_b200/multigpu.py), with CPU stand-in workers:
spawn, readiness, least-loaded dispatch by ViT tokens, error propagation, shutdown."""
import threading

import pytest
from PIL import Image

from dots_ocr_b200.multigpu import MultiGpuRunner, echo_worker


def test_requests_spread_over_workers_and_come_back_to_their_callers():
    r = MultiGpuRunner(2, factory=echo_worker, factory_args=(0.05,), start_timeout=120)
    try:
        small = Image.new("RGB", (1024, 1024))
        out = {}

        def one(i):
            out[i] = r.infer(small, f"p{i}", max_new_tokens=10 + i)
        ths = [threading.Thread(target=one, args=(i,)) for i in range(12)]
        [t.start() for t in ths]
        [t.join(60) for t in ths]
        assert len(out) == 12
        for i in range(12):
            rank, size, prompt, n = out[i].split("|")
            assert (size, prompt, n) == ("(1024, 1024)", f"p{i}", str(10 + i)) and rank in ("rank0", "rank1")
        assert sum(r.served) == 12 and min(r.served) >= 4       # equal pages -> (near) equal split
        assert r._load == [0, 0] and not r._pending
        # a 1960x1960 page weighs 3.6 small pages: the next three small pages all go to the other worker
        big = Image.new("RGB", (1960, 1960))
        futs = [r.submit(big, "big")] + [r.submit(small, f"s{k}") for k in range(3)]
        res = [f.result(timeout=60) for f in futs]
        assert res[0].startswith("rank0|") and all(x.startswith("rank1|") for x in res[1:])
        assert r.infer_batch([small, small], ["a", "b"], 7) == [f"rank{k}|(1024, 1024)|{p}|7" for k, p in ((0, "a"), (1, "b"))]
    finally:
        r.close()
    assert all(not p.is_alive() for p in r._procs)
    with pytest.raises(RuntimeError, match="closed"):
        r.submit("x", "y")


def test_worker_errors_reach_the_caller_and_a_dead_start_is_loud():
    r = MultiGpuRunner(2, factory=echo_worker, factory_args=(0.0, "boom"), start_timeout=120)
    try:
        assert r.infer("img", "fine").endswith("|img|fine|512")
        with pytest.raises(ValueError, match="refuses 'boom page'"):
            r.infer("img", "boom page")
        assert r.infer("img", "still fine").endswith("|still fine|512")     # the worker survives a failed page
    finally:
        r.close()
    with pytest.raises(RuntimeError, match="worker 1 failed to start"):
        MultiGpuRunner(2, factory=echo_worker, factory_args=(0.0, "", 1), start_timeout=120)


def test_a_dead_worker_fails_its_pages_and_the_rest_go_on():
    r = MultiGpuRunner(2, factory=echo_worker, factory_args=(0.0,), start_timeout=120)
    try:
        doomed = r.submit("img", "die now")                  # goes to worker 0 (both idle, lowest index)
        with pytest.raises(RuntimeError, match="GPU worker 0 exited with code 3"):
            doomed.result(timeout=30)
        outs = [r.infer("img", f"p{i}") for i in range(4)]   # every later page is served by the survivor
        assert all(o.startswith("rank1|") for o in outs)
    finally:
        r.close()
