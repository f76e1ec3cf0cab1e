"""Continuous batching on the device: the slot backend over the real engine (dots_ocr_b200/continuous.py:EngineSlots) must
give every page exactly the ids a one-page ``generate`` gives it, whatever shares the cache with it and whenever it was
admitted.  Tiny config, `random` checkpoint (the continuation depends on the whole context, so a row reading a neighbour's
or a previous tenant's keys would show).  (Passed on B200 in the round-1 driver run and in round 2's first GPU pass.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(flavour):
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    from dots_ocr_b200.processing import SyntheticTokenizer

    class IdTokenizer(SyntheticTokenizer):
        def decode(self, ids):
            ids = [int(i) for i in ids]
            if self.eos_token_id is not None and self.eos_token_id in ids:
                ids = ids[: ids.index(self.eos_token_id) + 1]          # keep the stop id, drop the pads after it
            return ",".join(str(i) for i in ids)

    cfg = config.tiny()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, flavour), DEV)
    return cfg, eng, IdTokenizer(cfg)


def _pages(n, seed=0):
    from PIL import Image
    rng = np.random.default_rng(seed)
    sizes = [(100, 80), (60, 120), (90, 90), (140, 70), (56, 56), (120, 120), (84, 168)]
    return [Image.fromarray(rng.integers(0, 256, (sizes[i % len(sizes)][1], sizes[i % len(sizes)][0], 3), dtype=np.uint8))
            for i in range(n)]


@pytest.mark.parametrize("flavour", ["random", "peaked"])
def test_slots_give_each_page_what_generate_gives_it(flavour):
    from dots_ocr_b200.continuous import ContinuousBatcher, EngineSlots
    from dots_ocr_b200.runner import PageRunner
    cfg, eng, tk = _setup(flavour)
    imgs = _pages(7)
    prompts = [f"page {i}: layout" for i in range(7)]
    budgets = [5, 20, 9, 13, 1, 24, 7]
    single = PageRunner(eng, tk)
    want = [single.infer_batch([im], [pr], max_new_tokens=b)[0] for im, pr, b in zip(imgs, prompts, budgets)]
    assert all(len(w.split(",")) == b for w, b in zip(want, budgets))

    backend = EngineSlots(eng, tk, n_slots=3, max_prompt=256, max_new=24, chunk=4)
    cb = ContinuousBatcher(backend)
    futs = [cb.submit(im, pr, b) for im, pr, b in zip(imgs, prompts, budgets)]
    got = [f.result(timeout=120) for f in futs]
    cb.close()
    assert got == want
    assert cb.stats["pages"] == 7 and cb.stats["max_active"] == 3 and cb.stats["admissions"] >= 3


def test_slots_stop_on_the_stop_id_and_reuse_the_row():
    from dots_ocr_b200 import weights
    from dots_ocr_b200.continuous import ContinuousBatcher, EngineSlots
    cfg, eng, tk = _setup("peaked")
    # peaked: the continuation is a known chain from the last prompt token (<|assistant|>)
    chain = [tk.assistant_id]
    for _ in range(12):
        chain.append(weights.peaked_next_token(cfg, chain[-1]))
    tk.eos_token_id = chain[6]                   # every page stops after 6 tokens (the 6th is the stop id itself)
    backend = EngineSlots(eng, tk, n_slots=2, max_prompt=256, max_new=16, chunk=3)
    cb = ContinuousBatcher(backend)
    imgs = _pages(5, seed=1)
    got = cb.infer_batch(imgs, ["p"] * 5, max_new_tokens=16)
    cb.close()
    want = ",".join(str(t) for t in chain[1:7])
    assert got == [want] * 5                     # 5 pages through 2 rows: each row was reused, nobody saw stale state
