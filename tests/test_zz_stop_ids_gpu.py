"""Stop ids and early exit of the decode loop on the device (tiny config, `peaked` checkpoint: the next token is a
known permutation of the previous one, so the step at which each row stops is chosen by the test).

The host halves (finalize_new_tokens vs HF generate, replay_steps) are pinned on CPU in tests/test_cpu_host.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def peaked():
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    cfg = config.tiny()
    ck = weights.make_synthetic_checkpoint(cfg, 0, "peaked")
    return cfg, Engine(cfg, ck, DEV)


def _chain(cfg, start, n):
    from dots_ocr_b200 import weights
    out = [start]
    for _ in range(n):
        out.append(weights.peaked_next_token(cfg, out[-1]))
    return out[1:]


def test_early_exit_gives_the_same_sequences_with_fewer_launches(peaked):
    cfg, eng = peaked
    ids = torch.tensor([[5, 6, 7, 8], [9, 10, 11, 12]])
    N = 96
    c0, c1 = _chain(cfg, 8, N), _chain(cfg, 12, N)
    eos = c0[20]
    assert eos not in c0[:20]
    # row 1 must also meet the stop id for the batch to end early: it does only if its chain crosses the same value
    stop1 = c1.index(eos) if eos in c1 else None
    old = eng.eos_check_every
    try:
        eng.eos_check_every = 0
        l0 = eng.launches
        full = eng.generate(ids, max_new_tokens=N, eos_token_id=eos, pad_token_id=0).sequences
        n_full = eng.launches - l0
        eng.eos_check_every = 8
        l0 = eng.launches
        early = eng.generate(ids, max_new_tokens=N, eos_token_id=eos, pad_token_id=0).sequences
        n_early = eng.launches - l0
    finally:
        eng.eos_check_every = old
    assert torch.equal(full, early)
    assert full[0, 4:4 + 21].tolist() == c0[:21] and (full[0, 4 + 21:] == 0).all()
    if stop1 is not None:
        assert full.shape[1] == 4 + max(20, stop1) + 1
        assert n_early < n_full
    else:
        assert full.shape[1] == 4 + N and n_early == n_full


def test_both_rows_stop_early_and_the_loop_leaves(peaked):
    """The same prompt in both rows: both stop at step 12, so with a look every 8 steps the loop leaves after 16 of 200."""
    cfg, eng = peaked
    ids = torch.tensor([[5, 6, 7, 8], [5, 6, 7, 8]])
    N = 200
    c = _chain(cfg, 8, 13)
    old = eng.eos_check_every
    try:
        eng.eos_check_every = 8
        l0 = eng.launches
        out = eng.generate(ids, max_new_tokens=N, eos_token_id=c[12], pad_token_id=0).sequences
        used = eng.launches - l0
        out2 = eng.generate(ids, max_new_tokens=N, eos_token_id=c[12], pad_token_id=0, use_graph=False).sequences
        assert torch.equal(out, out2)                       # the eager loop takes the same exit
    finally:
        eng.eos_check_every = old
    assert out.shape == (2, 4 + 13)
    assert out[0, 4:].tolist() == c and out[1, 4:].tolist() == c
    assert used < 30 * eng.launches_per_decode_step(2)          # ~17 steps, not 199


def test_secondary_stop_id_finishes_the_row_on_the_device(peaked):
    """generation_config.eos_token_id is a list and the model ends its turn with the SECOND id: the row's finished flag must
    still go up on the device, so the early exit fires and a continuous-batching slot is released."""
    cfg, eng = peaked
    ids = torch.tensor([[5, 6, 7, 8], [5, 6, 7, 8]])
    c = _chain(cfg, 8, 13)
    primary_never = cfg.text.vocab_size - 3
    assert primary_never not in c
    old = eng.eos_check_every
    try:
        eng.eos_check_every = 8
        l0 = eng.launches
        out = eng.generate(ids, max_new_tokens=200, eos_token_id=[primary_never, c[12]], pad_token_id=0).sequences
        used = eng.launches - l0
    finally:
        eng.eos_check_every = old
    assert out.shape == (2, 4 + 13) and out[0, 4:].tolist() == c
    assert used < 30 * eng.launches_per_decode_step(2)          # left after ~16 steps, not 199


def test_two_stop_ids_each_row_its_own(peaked):
    cfg, eng = peaked
    ids = torch.tensor([[5, 6, 7, 8], [9, 10, 11, 12]])
    N = 24
    c0, c1 = _chain(cfg, 8, N), _chain(cfg, 12, N)
    a, b = c0[5], c1[9]
    if a in c1[:9] or b in c0[:5]:
        pytest.skip("chains cross before the chosen stops")
    out = eng.generate(ids, max_new_tokens=N, eos_token_id=[a, b], pad_token_id=0).sequences[:, 4:]
    assert out.shape[1] == 10                                   # HF ends the loop when the last row stops (step 9)
    assert out[0].tolist() == c0[:6] + [0] * 4
    assert out[1].tolist() == c1[:10]
