"""Two batches in flight on one GPU (SURVEY.md section 8f "next": the serving loop above Engine.generate).

A batch's life has two phases with opposite bottlenecks: ViT encode + prefill keeps the tensor pipe (and the board's power
budget) busy and leaves HBM idle; the greedy decode loop streams weights and KV cache with the tensor pipe idle.  ``Engine.generate``
runs them back to back.  ``PagePipeline`` runs batch i+1's encode + prefill on one SM partition while batch i decodes on the other
(CUDA green contexts, csrc/partition.cu): two ordinary streams would not overlap, because the prefill kernels are persistent
one-CTA-per-SM grids that hold every SM for milliseconds at a time.

The arithmetic is the engine's: the same kernels with the same split plan, so a page's ids are those of ``Engine.generate`` on the
same batch (tests/test_pipeline_gpu.py).  Only the placement changes.

    pipe = PagePipeline(engine, prefill_sms=96)
    outs = pipe.run([dict(input_ids=ids, pixel_values=pv, image_grid_thw=grid, max_new_tokens=512), ...])
    pipe.close()
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .engine import Engine, GenerateOutput, _round_up, finalize_new_tokens, replay_steps, stop_list


class _Job:
    __slots__ = ("req", "ids", "B", "N", "ctx_max", "lens", "seq_lens", "ids_packed", "positions", "seq_of_tok", "cu", "last_rows",
                 "stops", "pad", "key", "ready", "prefilled", "out_new", "slot", "steps_done")


class PagePipeline:
    def __init__(self, eng: Engine, prefill_sms: int = 96, decode_plan_sms: int = 0):
        """``prefill_sms``: SMs (multiple of 8) given to encode + prefill; decode gets the rest.  ``decode_plan_sms`` = 0 keeps the
        whole-device split-K plan for the decode step (bit-identical ids to Engine.generate); n re-plans the splits for n SMs."""
        self.eng = eng
        with torch.cuda.device(eng.device):
            self.s_pre, self.s_dec, self.n_pre, self.n_dec = ops.partition(int(prefill_sms))
        self.decode_plan_sms = int(decode_plan_sms)
        self.slots: List[Optional[dict]] = [None, None]          # per slot: dict(key, st, graph, done event)
        self.closed = False

    # ------------------------------------------------------------------------------------------------------------------
    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        with torch.cuda.device(self.eng.device):
            torch.cuda.synchronize()
            self.slots = [None, None]          # KV caches, decode state, graphs.  The SM partition itself stays (ops.partition)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # ------------------------------------------------------------------------------------------------------------------
    def _admit(self, req: dict) -> _Job:
        """Host-side part of Engine.generate (packing, positions, cumulative lengths), done before anything is enqueued so that
        no later step has to look at the device."""
        eng, dev = self.eng, self.eng.device
        j = _Job()
        j.req = req
        ids_cpu = req["input_ids"].detach().to("cpu").long()
        B, Tpad = ids_cpu.shape
        assert B <= 64, "the pipeline runs the tiled decode step: at most 64 pages per batch"
        mask = req.get("attention_mask")
        mask = torch.ones_like(ids_cpu) if mask is None else mask.detach().to("cpu").long()
        lens = mask.sum(dim=1).to(torch.int64)
        j.B, j.N = B, int(req.get("max_new_tokens", 16))
        assert j.N >= 1
        j.seq_lens = lens.tolist()
        j.ctx_max = _round_up(int(max(j.seq_lens)) + j.N, 64)
        keep = mask.bool()
        packed = ids_cpu[keep].contiguous()
        n_img = int((packed == eng.cfg.image_token_id).sum())
        grid = req.get("image_grid_thw")
        if grid is not None:
            g = grid.tolist() if torch.is_tensor(grid) else grid
            want = sum(int(t) * int(h) * int(w) for t, h, w in g) // (eng.cfg.vision.spatial_merge_size ** 2)
            if n_img != want:
                raise ValueError(f"image tokens in input_ids ({n_img}) != image embedding rows ({want})")
        positions = (mask.cumsum(1) - 1)[keep].to(torch.int32).contiguous()
        seq_of_tok = torch.arange(B, dtype=torch.int32).unsqueeze(1).expand(B, Tpad)[keep].contiguous()
        cu = torch.zeros(B + 1, dtype=torch.int32)
        cu[1:] = lens.cumsum(0).to(torch.int32)
        up = lambda t: t.pin_memory().to(dev, non_blocking=True)
        j.ids = up(ids_cpu)
        j.ids_packed, j.positions, j.seq_of_tok, j.cu = up(packed), up(positions), up(seq_of_tok), up(cu)
        j.lens = up(lens)
        j.last_rows = up((cu[1:] - 1).to(torch.int32))
        j.stops = stop_list(req.get("eos_token_id"))
        j.pad = int(req.get("pad_token_id", 0))
        j.key = (B, j.ctx_max, j.N, tuple(j.stops[: ops.MAX_STOP_IDS]), j.pad, self.decode_plan_sms)
        j.ready = torch.cuda.Event()
        j.ready.record()
        j.prefilled = torch.cuda.Event()
        j.out_new = None
        j.steps_done = 0
        return j

    def _slot_for(self, j: _Job, s: int) -> dict:
        ent = self.slots[s]
        if ent is not None and ent["key"] == j.key:
            return ent
        if ent is not None:
            ent["done"].synchronize()                 # the slot's previous batch must be out of its buffers before they are released
            self.slots[s] = ent = None
        eng = self.eng
        eng.decode_sms = self.decode_plan_sms
        try:
            kc, vc = eng._alloc_cache(j.B, j.ctx_max)
            st = eng._new_decode_state(j.B, j.lens, kc, vc, j.ctx_max, j.N, j.stops or None, j.pad)
        finally:
            eng.decode_sms = 0
        ent = dict(key=j.key, st=st, graph=None, done=torch.cuda.Event(), fresh=True)
        self.slots[s] = ent
        return ent

    def _prefill(self, j: _Job, ent: dict) -> None:
        """Encode + prefill + first token of one batch, enqueued on the prefill partition (no host synchronisation inside)."""
        eng, t, req = self.eng, self.eng.cfg.text, j.req
        st = ent["st"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ent["done"])                    # the batch that used this slot before has left its KV cache and state
        if not ent.pop("fresh", False):
            st["last"].zero_()
            st["out_ids"].fill_(j.pad)
            st["step"].zero_()
            st["finished"].zero_()
            st["pos"].copy_((j.lens - 1).to(torch.int32))
            st["ctx_len"].copy_(j.lens.to(torch.int32))
        image_embeds = req.get("image_embeds")
        if image_embeds is None and req.get("pages_u8") is not None:
            pages = [p.to(eng.device, non_blocking=True) for p in req["pages_u8"]]
            image_embeds = eng.encode_pages_u8(pages, min_pixels=req.get("min_pixels"), max_pixels=req.get("max_pixels"))
        elif image_embeds is None and req.get("pixel_values") is not None:
            image_embeds = eng.encode_images(req["pixel_values"].to(eng.device, non_blocking=True), req["image_grid_thw"])
        slots = None
        if image_embeds is not None:
            slots, _count = ops.image_slots(j.ids_packed, eng.cfg.image_token_id)
            eng.launches += 1
        x = eng._prefill(j.ids_packed, slots, image_embeds, j.cu, j.seq_lens, j.positions, j.seq_of_tok, st["kc"], st["vc"], j.ctx_max)
        hl = ops.gather_rows(x, j.last_rows)
        ops.rmsnorm(hl, eng.final_norm, t.rms_norm_eps, out=st["normed"])
        ops.gemm_skinny(st["normed"], eng.lm_head, 1, out_bf16=st["logits"])
        ops.argmax_advance(st["logits"], st["last"], st["out_ids"], st["step"], st["pos"], st["ctx_len"], st["finished"],
                           st["stops"], st["pad"], st["forced"])
        eng.launches += 4
        j.prefilled.record()

    def _decode(self, j: _Job, ent: dict) -> None:
        """The decode loop of one batch on the decode partition: the captured step, replayed."""
        eng, st = self.eng, ent["st"]
        cur = torch.cuda.current_stream()
        cur.wait_event(j.prefilled)
        n_steps = j.N - 1
        done = 0
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        if n_steps > 0:
            if ent["graph"] is None:
                eng._decode_step(st)                   # eager once (also warms every kernel variant), then capture on this very stream
                done = 1
                if n_steps > done:
                    ent["graph"] = ops.capture(lambda: eng._decode_step(st))
            if n_steps > done:
                check = int(eng.eos_check_every) if st["stops"] else 0
                done += replay_steps(ent["graph"].launch, n_steps - done, check, lambda: bool(st["finished"].all().item()))
            eng.launches += eng.launches_per_decode_step(j.B) * done
        ev[1].record()
        eng.decode_log.append((eng.decode_bytes(j.B, j.seq_lens, done), done, ev[0], ev[1]))
        del eng.decode_log[:-64]
        j.steps_done = done
        j.out_new = finalize_new_tokens(st["out_ids"], j.stops, j.pad).clone()
        ent["done"].record()

    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, requests: List[dict]) -> List[GenerateOutput]:
        """Generate for every request (the keyword arguments of Engine.generate as a dict); request i+1 is encoded and prefilled
        while request i decodes.  Host tensors should be pinned: a pageable host->device copy can block the enqueueing thread."""
        assert not self.closed
        eng = self.eng
        with torch.cuda.device(eng.device):
            jobs = [self._admit(r) for r in requests]
            ents: List[Optional[dict]] = [None] * len(jobs)
            for i in range(len(jobs) + 1):
                if i < len(jobs):
                    with ops.on_partition(self.s_pre, self.n_pre):
                        torch.cuda.current_stream().wait_event(jobs[i].ready)       # the admission uploads (made on the caller's stream)
                        ents[i] = self._slot_for(jobs[i], i % 2)
                        self._prefill(jobs[i], ents[i])
                if i >= 1:
                    eng.decode_sms = self.decode_plan_sms
                    try:
                        with ops.on_partition(self.s_dec, self.n_dec):
                            self._decode(jobs[i - 1], ents[i - 1])
                    finally:
                        eng.decode_sms = 0
            outs = []
            main = torch.cuda.current_stream()
            for s in self.slots:
                if s is not None:
                    main.wait_event(s["done"])
            for j in jobs:
                j.out_new.record_stream(main)          # allocated on the decode partition's stream, read on the caller's
                outs.append(GenerateOutput(sequences=torch.cat([j.ids, j.out_new], dim=1)))
            return outs
