"""PagePipeline (dots_ocr_b200/pipeline.py): two batches in flight on SM partitions give every page the ids of Engine.generate on the
same batch -- different batches (ragged prompts, different page sizes, uint8 pages), more batches than slots (slot reuse), stop ids."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(cfg, seed, grids):
    g = torch.Generator().manual_seed(seed)
    pvs, rows = [], []
    for (_, h, w) in grids:
        pvs.append(torch.randn(h * w, cfg.vision.patch_dim, generator=g))
        rows.append(torch.cat([torch.randint(0, 2000, (3 + seed % 4,), generator=g), torch.full((h * w // 4,), cfg.image_token_id),
                               torch.randint(0, 2000, (5 + seed % 3,), generator=g)]))
    T = max(r.numel() for r in rows)
    ids = torch.zeros((len(rows), T), dtype=torch.long)
    mask = torch.zeros((len(rows), T), dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, T - r.numel():] = r
        mask[i, T - r.numel():] = 1
    return dict(input_ids=ids, attention_mask=mask, pixel_values=torch.cat(pvs).pin_memory(), image_grid_thw=torch.tensor(grids))


@pytest.mark.parametrize("flavour", ["random", "peaked"])
def test_pipeline_matches_generate(flavour):
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    from dots_ocr_b200.pipeline import PagePipeline
    cfg = config.tiny()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, flavour), DEV)
    shapes = [[(1, 8, 8), (1, 4, 12), (1, 8, 12)], [(1, 8, 8), (1, 4, 12), (1, 8, 12)], [(1, 12, 8), (1, 8, 8)],
              [(1, 8, 8), (1, 4, 12), (1, 8, 12)], [(1, 8, 8), (1, 4, 12), (1, 8, 12)]]
    reqs = []
    for i, gr in enumerate(shapes):
        r = _batch(cfg, 11 + i, gr)
        r["max_new_tokens"] = 24 if i != 2 else 17
        reqs.append(r)
    want = [eng.generate(r["input_ids"], attention_mask=r["attention_mask"], pixel_values=r["pixel_values"].to(DEV),
                         image_grid_thw=r["image_grid_thw"], max_new_tokens=r["max_new_tokens"]).sequences.cpu() for r in reqs]
    with PagePipeline(eng, prefill_sms=96) as pipe:
        got = [o.sequences.cpu() for o in pipe.run(reqs)]
        again = [o.sequences.cpu() for o in pipe.run(reqs[:3])]             # slots and graphs are reused across calls
    for i, (a, b) in enumerate(zip(want, got)):
        assert torch.equal(a, b), f"batch {i} differs"
    for a, b in zip(want[:3], again):
        assert torch.equal(a, b)
    # the engine is still usable on the whole device afterwards
    r = reqs[0]
    after = eng.generate(r["input_ids"], attention_mask=r["attention_mask"], pixel_values=r["pixel_values"].to(DEV),
                         image_grid_thw=r["image_grid_thw"], max_new_tokens=r["max_new_tokens"]).sequences.cpu()
    assert torch.equal(after, want[0])


def test_pipeline_stop_ids_and_u8_pages():
    from dots_ocr_b200 import config, weights
    from dots_ocr_b200.engine import Engine
    from dots_ocr_b200.pipeline import PagePipeline
    cfg = config.tiny()
    eng = Engine(cfg, weights.make_synthetic_checkpoint(cfg, 0, "peaked"), DEV)
    ids = torch.tensor([[5, 6, 7, 8], [9, 10, 11, 12]])
    chain = [8]
    for _ in range(40):
        chain.append(weights.peaked_next_token(cfg, chain[-1]))
    eos = chain[13]
    reqs = [dict(input_ids=ids, max_new_tokens=48, eos_token_id=[eos, 3], pad_token_id=0),
            dict(input_ids=ids + 1, max_new_tokens=48, eos_token_id=[eos, 3], pad_token_id=0),
            dict(input_ids=ids, max_new_tokens=48, eos_token_id=[eos, 3], pad_token_id=0)]
    want = [eng.generate(r["input_ids"], max_new_tokens=48, eos_token_id=[eos, 3], pad_token_id=0).sequences.cpu() for r in reqs]
    g = torch.Generator().manual_seed(3)
    f = cfg.vision.patch_size * cfg.vision.spatial_merge_size
    page = torch.randint(0, 256, (4 * f + 5, 6 * f - 3, 3), generator=g, dtype=torch.uint8)
    probe = eng.encode_pages_u8([page.to(DEV)])
    n_img = probe.shape[0]
    row = torch.cat([torch.tensor([1, 2, 3]), torch.full((n_img,), cfg.image_token_id), torch.tensor([4, 5])]).unsqueeze(0)
    u8 = dict(input_ids=row, pages_u8=[page.pin_memory()], max_new_tokens=12)
    want_u8 = eng.generate(row, pages_u8=[page.to(DEV)], max_new_tokens=12).sequences.cpu()
    with PagePipeline(eng, prefill_sms=96) as pipe:
        got = [o.sequences.cpu() for o in pipe.run(reqs)]
        got_u8 = pipe.run([u8, u8])
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    assert torch.equal(got_u8[0].sequences.cpu(), want_u8) and torch.equal(got_u8[1].sequences.cpu(), want_u8)
