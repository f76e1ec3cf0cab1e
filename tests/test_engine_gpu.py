"""End-to-end parity of the CUDA engine against the oracle (tiny config; same seeded weights).

* vision tower: per-layer activations vs the fp32 CPU oracle (T0) and the bf16 CUDA oracle (T1)
* greedy ids: bit-exact vs HF generate on the `peaked` checkpoint (well-posed argmax margins)
* pre-sampling logits: teacher-forced, `random` checkpoint, tolerance stated below
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Tolerances for bf16 pre-sampling logits, in units of std(oracle logits) (max over batch, step, vocab):
#  * vs the reference HF forward in bf16 on the same GPU (T1): <= LOGIT_TOL_T1.  bf16 logits of
#    magnitude 2-4 sigma are quantised in steps of ~0.03 sigma, so two correct bf16 pipelines differ by
#    1-2 such steps at the worst element (measured: 0.034-0.041); real bugs show up as O(1).
#  * vs the fp32 CPU oracle (T0): no worse than 1.5x what HF's own bf16 forward shows against T0
#    (measured 0.029-0.039 for both), floor LOGIT_TOL_T0_FLOOR.
LOGIT_TOL_T1 = 6e-2
LOGIT_TOL_T0_FLOOR = 4e-2


def _inputs(cfg, grids, n_text_front=5, n_text_back=7, seed=7):
    g = torch.Generator().manual_seed(seed)
    pvs, ids_rows = [], []
    for (t, h, w) in grids:
        S = t * h * w
        pvs.append(torch.randn(S, cfg.vision.patch_dim, generator=g))
        row = torch.cat([torch.randint(0, cfg.text.vocab_size - 16, (n_text_front,), generator=g),
                         torch.full((S // 4,), cfg.image_token_id),
                         torch.randint(0, cfg.text.vocab_size - 16, (n_text_back,), generator=g)])
        ids_rows.append(row)
    return torch.cat(pvs), torch.tensor(grids), ids_rows


@pytest.fixture(scope="module")
def tiny():
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    cfg = config.tiny()
    out = {}
    for fl in ("peaked", "random"):
        ck = weights.make_synthetic_checkpoint(cfg, 0, fl)
        out[fl] = (ck, Engine(cfg, ck, DEV))
    return cfg, out


def test_vision_tower_layers(tiny):
    from oracle.vision import VisionOracle
    cfg, d = tiny
    ck, eng = d["random"]
    pv, grid, _ = _inputs(cfg, [(1, 8, 12), (1, 6, 6), (1, 16, 16)])
    out, layers = eng.encode_images(pv.to(DEV), grid, return_layers=True)
    o32 = VisionOracle(cfg.vision, ck, torch.float32, "cpu")
    ref, ref_layers = o32.forward(pv, grid, return_layers=True)
    for i, (a, b) in enumerate(zip(layers, ref_layers)):
        err = float((a.float().cpu() - b).abs().max() / b.abs().max())
        assert err < 3e-2, (i, err)
    err = float((out.float().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 3e-2, err
    # T1: same arithmetic in bf16 on the GPU through plain torch ops
    o16 = VisionOracle(cfg.vision, ck, torch.bfloat16, DEV)
    ref16 = o16.forward(pv.to(DEV), grid)
    err16 = float((out.float() - ref16.float()).abs().max() / ref16.float().abs().max())
    assert err16 < 3e-2, err16


def test_greedy_ids_bit_exact_peaked(tiny):
    from oracle.model import DotsOracle
    cfg, d = tiny
    ck, eng = d["peaked"]
    grids = [(1, 8, 8), (1, 8, 8)]
    pv, grid, rows = _inputs(cfg, grids)
    ids = torch.stack(rows)
    N = 24
    ref = DotsOracle(cfg, ck, torch.float32, "cpu").generate(ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N)
    ref16 = DotsOracle(cfg, ck, torch.bfloat16, DEV).generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N)
    got = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N).sequences
    assert got.shape == ref.shape
    assert torch.equal(got.cpu(), ref), (got[:, -N:].tolist(), ref[:, -N:].tolist())
    assert torch.equal(got, ref16.to(got.device))
    # graph replay and eager stepping agree
    got2 = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N, use_graph=False).sequences
    assert torch.equal(got, got2)


def test_ragged_batch_left_padded(tiny):
    """HF batches are left-padded with an attention mask (parser.py:99-105 `padding=True`)."""
    from oracle.model import DotsOracle
    cfg, d = tiny
    ck, eng = d["peaked"]
    pv, grid, rows = _inputs(cfg, [(1, 8, 8), (1, 4, 12)])
    T = max(r.numel() for r in rows)
    ids = torch.zeros((2, T), dtype=torch.long)
    mask = torch.zeros((2, T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    N = 12
    ref = DotsOracle(cfg, ck, torch.float32, "cpu").generate(ids, attention_mask=mask, pixel_values=pv, image_grid_thw=grid,
                                                              max_new_tokens=N)
    got = eng.generate(ids, attention_mask=mask, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N).sequences
    assert torch.equal(got.cpu(), ref)


def test_teacher_forced_logits_random(tiny):
    from oracle.model import DotsOracle
    cfg, d = tiny
    ck, eng = d["random"]
    pv, grid, rows = _inputs(cfg, [(1, 8, 8), (1, 8, 8), (1, 8, 8)])
    ids = torch.stack(rows)
    N = 16
    o32 = DotsOracle(cfg, ck, torch.float32, "cpu")
    ref_ids = o32.generate(ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N)
    new = ref_ids[:, ids.shape[1]:]
    ref_logits = o32.teacher_forced_logits(ids, new, pv, grid)                       # [B, N, V] fp32
    out = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N, forced_ids=new, return_logits=True)
    assert torch.equal(out.sequences[:, ids.shape[1]:].cpu(), new)                   # forcing took effect
    got = out.logits.float().cpu()
    sd = ref_logits.std()
    ref16 = DotsOracle(cfg, ck, torch.bfloat16, DEV).teacher_forced_logits(ids, new, pv.to(DEV), grid).cpu()
    err_t1 = float((got - ref16).abs().max() / sd)
    err_t0 = float((got - ref_logits).abs().max() / sd)
    hf_t0 = float((ref16 - ref_logits).abs().max() / sd)
    print(f"logit err / sigma: engine-vs-HF-bf16 {err_t1:.4f}, engine-vs-fp32 {err_t0:.4f}, HF-bf16-vs-fp32 {hf_t0:.4f}")
    assert err_t1 < LOGIT_TOL_T1, err_t1
    assert err_t0 < max(1.5 * hf_t0, LOGIT_TOL_T0_FLOOR), (err_t0, hf_t0)
    # wherever the oracle's top-1 margin clears the tolerance, the engine's argmax must agree
    top2 = ref_logits.topk(2, -1).values
    clear = (top2[..., 0] - top2[..., 1]) > 2 * LOGIT_TOL_T1 * sd
    assert clear.float().mean() > 0.3
    assert torch.equal(got.argmax(-1)[clear], ref_logits.argmax(-1)[clear])


def test_eos_and_pad(tiny):
    from dots_ocr_b200 import weights
    cfg, d = tiny
    ck, eng = d["peaked"]
    ids = torch.tensor([[5, 6, 7, 8], [9, 10, 11, 12]])
    # choose the 3rd token row 0 will emit as EOS: row 0 then pads, row 1 keeps going
    chain = [8]
    for _ in range(3):
        chain.append(weights.peaked_next_token(cfg, chain[-1]))
    eos = chain[3]
    out = eng.generate(ids, max_new_tokens=8, eos_token_id=eos, pad_token_id=0).sequences[:, 4:]
    assert out[0, :3].tolist() == chain[1:4] and out[0, 3:].tolist() == [0] * 5
    assert (out[1] != 0).all()


def test_pdl_on_off_identical(tiny):
    """Programmatic dependent launch only changes when kernels start, never what they compute."""
    from dots_ocr_b200 import ops
    cfg, d = tiny
    ck, eng = d["random"]
    pv, grid, rows = _inputs(cfg, [(1, 8, 8), (1, 8, 12)], n_text_back=11)
    T = max(r.numel() for r in rows)
    ids = torch.zeros((2, T), dtype=torch.long)
    mask = torch.zeros((2, T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    outs = []
    try:
        for flag in (False, True, True):
            ops.set_pdl(flag)
            o = eng.generate(ids, attention_mask=mask, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=40)
            outs.append(o.sequences.cpu())
    finally:
        ops.set_pdl(True)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


def test_gpu_preprocessing_path_matches_host(tiny):
    """uint8 pages normalised / patchified on the GPU (Engine.generate(pages_u8=...)) == host processor output fed as pixel_values."""
    from dots_ocr_b200.processing import preprocess_image
    cfg, d = tiny
    ck, eng = d["peaked"]
    g = torch.Generator().manual_seed(21)
    imgs = [torch.randint(0, 256, (112, 168, 3), generator=g, dtype=torch.uint8), torch.randint(0, 256, (224, 112, 3), generator=g, dtype=torch.uint8)]
    pvs, grids, rows = [], [], []
    for im in imgs:
        pv, gr = preprocess_image(im.numpy(), min_pixels=im.shape[0] * im.shape[1], max_pixels=im.shape[0] * im.shape[1])
        pvs.append(pv); grids.append(gr)
        rows.append(torch.cat([torch.randint(0, 2000, (4,), generator=g), torch.full((pv.shape[0] // 4,), cfg.image_token_id),
                               torch.randint(0, 2000, (3,), generator=g)]))
    T = max(r.numel() for r in rows)
    ids = torch.zeros((2, T), dtype=torch.long)
    mask = torch.zeros((2, T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    a = eng.generate(ids, attention_mask=mask, pixel_values=torch.cat(pvs).to(DEV), image_grid_thw=torch.cat(grids), max_new_tokens=10)
    b = eng.generate(ids, attention_mask=mask, pages_u8=[im.to(DEV) for im in imgs], max_new_tokens=10)
    assert torch.equal(a.image_embeds, b.image_embeds)
    assert torch.equal(a.sequences, b.sequences)


def test_batching_runner_matches_single_page_calls(tiny):
    """The reference parser's fan-out (threads, one page per call) through BatchingRunner == the same pages one by one."""
    import threading
    from PIL import Image
    from dots_ocr_b200.batching import BatchingRunner
    from dots_ocr_b200.processing import SyntheticTokenizer
    from dots_ocr_b200.runner import PageRunner
    cfg, d = tiny
    ck, eng = d["peaked"]
    runner = PageRunner(eng, SyntheticTokenizer(cfg))
    g = torch.Generator().manual_seed(77)
    sizes = [(112, 168), (224, 112), (140, 140), (56, 280), (168, 168), (112, 112)]
    pages = [Image.fromarray(torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).numpy()) for h, w in sizes]
    prompts = [f"prompt {i} " + "x" * i for i in range(len(pages))]
    single = [runner.infer(pg, pr, max_new_tokens=12) for pg, pr in zip(pages, prompts)]
    br = BatchingRunner(runner, max_batch=8, max_wait_ms=200)
    out = [None] * len(pages)

    def work(i):
        out[i] = br.infer(pages[i], prompts[i], max_new_tokens=12)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(pages))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    br.close()
    assert out == single
    assert sum(br.batches) == len(pages) and len(br.batches) < len(pages)          # at least some calls shared a generate


def test_parity_checks_are_sensitive_to_attention(tiny):
    """the `peaked` id test cannot see the attention (the successor is a function of the last
    token).  The `random`-weight checks can: with the test-only fault switch on (decode attention loses the P*V term of its
    first 64-key tile -- at these context lengths that is every key) the teacher-forced logits criterion of
    test_teacher_forced_logits_random fails by more than 10x and the greedy ids change, while the peaked ids do not move."""
    from dots_ocr_b200 import ops
    from oracle.model import DotsOracle
    cfg, d = tiny
    ck, eng = d["random"]
    pv, grid, rows = _inputs(cfg, [(1, 8, 8), (1, 8, 8)])
    ids = torch.stack(rows)
    N = 12
    o32 = DotsOracle(cfg, ck, torch.float32, "cpu")
    new = o32.generate(ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=N)[:, ids.shape[1]:]
    ref_logits = o32.teacher_forced_logits(ids, new, pv, grid)
    sd = ref_logits.std()
    res = {}
    try:
        for code in (0, 1):
            ops.debug_set_fault(code)
            out = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N, forced_ids=new, return_logits=True)
            free = eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N).sequences[:, ids.shape[1]:].cpu()
            res[code] = (float((out.logits.float().cpu()[:, 1:] - ref_logits[:, 1:]).abs().max() / sd), free)
        pk_ck, pk_eng = d["peaked"]
        pk = {}
        for code in (0, 1):
            ops.debug_set_fault(code)
            pk[code] = pk_eng.generate(ids, pixel_values=pv.to(DEV), image_grid_thw=grid, max_new_tokens=N).sequences.cpu()
    finally:
        ops.debug_set_fault(0)
    print(f"decode-step logits error / sigma: clean {res[0][0]:.4f}, with the fault {res[1][0]:.4f}")
    assert res[0][0] < LOGIT_TOL_T0_FLOOR * 1.5
    assert res[1][0] > 10 * LOGIT_TOL_T1, res[1][0]
    assert not torch.equal(res[0][1], res[1][1]), "greedy ids on random weights must react to broken attention"
    assert torch.equal(pk[0], pk[1])      # ... which is exactly why the peaked id test is plumbing only
