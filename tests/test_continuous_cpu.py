"""Scheduler of the continuous batcher (dots_ocr_b200/continuous.py) against a simulated slot backend that follows the
device kernels' state machine (argmax_advance: every row -- finished or idle too -- advances step / pos / ctx_len each
step and writes out_ids[row, step]) and ASSERTS the row bounds, so a scheduler that forgets to park idle rows or harvests
late fails here instead of corrupting a KV cache on the GPU."""
import threading
import time

import pytest

from dots_ocr_b200.continuous import ContinuousBatcher

EOS, PAD = 7, 0


class SimSlots:
    """Prompt "n=<k>" makes the row emit tokens 100, 101, ... (k of them) and then EOS."""

    def __init__(self, n_slots=3, max_prompt=16, max_new=40, chunk=4, step_sleep=0.0):
        self.n_slots, self.max_prompt, self.max_new, self.chunk = n_slots, max_prompt, max_new, chunk
        self.n_cols = max_new + chunk
        self.ctx_max = max_prompt + max_new + chunk
        self.step_sleep = step_sleep
        z = lambda v: [v] * n_slots
        self.stepc, self.pos, self.ctx, self.fin = z(0), z(0), z(1), z(1)
        self.out = [[PAD] * self.n_cols for _ in range(n_slots)]
        self.natural = z(0)
        self.log = []                       # ("admit", slots) / ("step", k, active rows)
        self.fail_admit = False

    def prepare(self, image, prompt):
        if "bad" in prompt:
            raise ValueError("cannot preprocess this page")
        k = int(prompt.split("=")[1])
        return (image, [1] * 5 + [k])      # (page, prompt ids); the last id carries the natural length

    def _emit(self, s):
        assert self.stepc[s] < self.n_cols, f"row {s}: out_ids overflow at step {self.stepc[s]}"
        assert self.pos[s] < self.ctx_max, f"row {s}: KV position {self.pos[s]} outside the cache row"
        if self.fin[s]:
            tok = PAD
        else:
            tok = 100 + self.stepc[s] if self.stepc[s] < self.natural[s] else EOS
            if tok == EOS:
                self.fin[s] = 1
        self.out[s][self.stepc[s]] = tok
        self.stepc[s] += 1
        self.pos[s] += 1
        self.ctx[s] += 1

    def admit(self, slots, prepared):
        if self.fail_admit:
            raise RuntimeError("prefill failed")
        self.log.append(("admit", list(slots)))
        for s, (_, ids) in zip(slots, prepared):
            assert self.fin[s] == 1 and self.stepc[s] == 0, f"slot {s} was not parked before admission"
            T = len(ids)
            self.natural[s] = ids[-1]
            self.out[s] = [PAD] * self.n_cols
            self.stepc[s], self.pos[s], self.ctx[s], self.fin[s] = 0, T - 1, T, 0
            self._emit(s)                   # first token comes out of the prefill

    def step(self, k):
        self.log.append(("step", k, [s for s in range(self.n_slots) if not self.fin[s]]))
        for _ in range(k):
            for s in range(self.n_slots):
                self._emit(s)
        if self.step_sleep:
            time.sleep(self.step_sleep)

    def poll(self):
        return list(self.stepc), list(self.fin)

    def take(self, slot, n):
        return self.out[slot][:n]

    def rearm(self, slots):
        for s in slots:
            self.stepc[s], self.pos[s], self.ctx[s], self.fin[s] = 0, 0, 1, 1

    def decode_text(self, ids):
        ids = ids[: ids.index(EOS)] if EOS in ids else ids
        return ",".join(str(i) for i in ids)


def _want(k, budget=10 ** 9):
    return ",".join(str(100 + i) for i in range(min(k, budget)))


def test_rows_are_refilled_while_others_keep_decoding():
    sim = SimSlots(n_slots=3, chunk=4, step_sleep=0.002)
    cb = ContinuousBatcher(sim)
    lengths = [3, 30, 5, 2, 9, 1, 14, 6, 0, 11]
    futs = [cb.submit(f"img{i}", f"n={k}") for i, k in enumerate(lengths)]
    got = [f.result(timeout=30) for f in futs]
    cb.close()
    assert got == [_want(k) for k in lengths]
    assert cb.stats["pages"] == 10 and cb.stats["max_active"] == 3
    # the 30-token page occupied its slot across several admissions: the others came and went beside it
    admits = [e for e in sim.log if e[0] == "admit"]
    assert len(admits) >= 4 and admits[0][1] == [0, 1, 2]
    # far fewer decode steps than one batch after the other would need: (3 pages per batch, each run to its longest)
    steps = sum(e[1] for e in sim.log if e[0] == "step")
    run_to_longest = sum(max(lengths[i:i + 3]) + 1 for i in range(0, len(lengths), 3))
    assert steps < run_to_longest, (steps, run_to_longest)


def test_budgets_stop_rows_and_idle_rows_stay_inside_their_cache_row():
    sim = SimSlots(n_slots=2, max_new=12, chunk=5)
    cb = ContinuousBatcher(sim)
    a = cb.submit("a", "n=100", max_new_tokens=7)          # natural length 100, budget 7
    b = cb.submit("b", "n=100", max_new_tokens=999)        # clamped to the session's max_new (12)
    assert a.result(timeout=30) == _want(100, 7) and b.result(timeout=30) == _want(100, 12)
    # one page at a time for a while: the second slot idles through many chunks (the simulator asserts its bounds)
    for i in range(6):
        assert cb.infer(f"x{i}", "n=11", max_new_tokens=12) == _want(11)
    cb.close()
    assert cb.stats["pages"] == 8


def test_errors_belong_to_their_pages():
    sim = SimSlots(n_slots=2, chunk=3)
    cb = ContinuousBatcher(sim)
    good, bad = cb.submit("g", "n=4"), cb.submit("b", "bad n=4")
    with pytest.raises(ValueError, match="cannot preprocess"):
        bad.result(timeout=30)
    assert good.result(timeout=30) == _want(4)
    sim.fail_admit = True
    f1, f2 = cb.submit("p", "n=2"), cb.submit("q", "n=3")
    for f in (f1, f2):
        with pytest.raises(RuntimeError, match="prefill failed"):
            f.result(timeout=30)
    sim.fail_admit = False
    assert cb.infer("r", "n=5") == _want(5)                 # the loop survived
    cb.close()
    with pytest.raises(RuntimeError, match="closed"):
        cb.submit("s", "n=1")


def test_admission_respects_the_vit_token_budget_and_close_drains():
    from PIL import Image
    sim = SimSlots(n_slots=4, chunk=2, step_sleep=0.01)
    cb = ContinuousBatcher(sim, max_admit_tokens=12000)     # two 1024x1024 pages (5476 each) per admission, not three
    page = Image.new("RGB", (1024, 1024))
    futs = [cb.submit(page, f"n={k}") for k in (20, 20, 20, 20)]
    cb.close(timeout=60)                                    # everything submitted before close is still answered
    assert [f.result(timeout=1) for f in futs] == [_want(20)] * 4
    admits = [e[1] for e in sim.log if e[0] == "admit"]
    assert admits[0] == [0, 1] and admits[1] == [2, 3]


def test_many_threads_one_page_each():
    sim = SimSlots(n_slots=8, max_new=40, chunk=4)
    cb = ContinuousBatcher(sim)
    out = {}

    def one(i):
        out[i] = cb.infer(f"img{i}", f"n={i % 13}", max_new_tokens=40)
    ths = [threading.Thread(target=one, args=(i,)) for i in range(64)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
    cb.close()
    assert len(out) == 64 and all(out[i] == _want(i % 13) for i in range(64))
    assert cb.stats["max_active"] <= 8
