"""Real-checkpoint plumbing on CPU: config.json -> DotsConfig, safetensors directory -> validated tensors,
tokenizer directory -> HFTokenizer, and PageRunner.from_checkpoint wiring them (with a recording stand-in for the
engine: no kernels run here).  No dots.ocr weights exist offline, so the directory is written by the test."""
import json
import os

import pytest
import torch

from dots_ocr_b200 import config as C
from dots_ocr_b200 import weights as W
from dots_ocr_b200.processing import HFTokenizer, build_text_inputs

SPECIALS = ["<|endoftext|>", "<|user|>", "<|endofuser|>", "<|assistant|>", "<|endofassistant|>", "<|img|>", "<|endofimg|>",
            "<|imgpad|>"]


def _tiny_hf_config(cfg):
    t, v = cfg.text, cfg.vision
    return dict(model_type="dots_ocr", hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
                num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
                num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size, rms_norm_eps=t.rms_norm_eps,
                rope_theta=t.rope_theta, max_position_embeddings=t.max_position_embeddings, tie_word_embeddings=False,
                hidden_act="silu", image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
                vision_config=dict(embed_dim=v.embed_dim, hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                                   num_hidden_layers=v.num_hidden_layers, num_attention_heads=v.num_attention_heads,
                                   num_channels=3, patch_size=14, spatial_merge_size=2, temporal_patch_size=1,
                                   rms_norm_eps=v.rms_norm_eps, use_bias=False, post_norm=True, is_causal=False,
                                   attn_implementation="flash_attention_2"))


def _write_tokenizer(path, chat_template=None):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {c: i for i, c in enumerate(sorted(pre_tokenizers.ByteLevel.alphabet()))}
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token=SPECIALS[0], pad_token=SPECIALS[0],
                                   additional_special_tokens=SPECIALS[1:])
    if chat_template:
        fast.chat_template = chat_template
    fast.save_pretrained(path)
    return fast


def test_config_json_defaults_are_the_published_architecture():
    assert C.from_hf_dict({}) == C.full()
    d = _tiny_hf_config(C.tiny())
    got = C.from_hf_dict(d)
    assert got.text == C.tiny().text and got.vision == C.tiny().vision and got.image_token_id == C.tiny().image_token_id
    # transformers >= 5 moves rope_theta under rope_parameters
    d2 = dict(d)
    d2.pop("rope_theta")
    d2["rope_parameters"] = {"rope_theta": 5.0e5, "rope_type": "default"}
    assert C.from_hf_dict(d2).text.rope_theta == 5.0e5


@pytest.mark.parametrize("patch,needle", [
    ({"tie_word_embeddings": True}, "tie_word_embeddings"),
    ({"use_sliding_window": True}, "sliding"),
    ({"rope_scaling": {"rope_type": "yarn", "factor": 4.0}}, "rope_scaling"),
    ({"num_attention_heads": 16}, "head_dim"),
    ({"vision_config": {"use_bias": True}}, "use_bias"),
    ({"vision_config": {"patch_size": 16}}, "patch geometry"),
    ({"hidden_act": "gelu"}, "hidden_act"),
])
def test_config_json_refuses_what_the_kernels_do_not_implement(patch, needle):
    with pytest.raises(C.UnsupportedCheckpoint, match=needle):
        C.from_hf_dict(patch)


def test_safetensors_directory_round_trip_and_validation(tmp_path):
    cfg = C.tiny()
    ckpt = W.make_synthetic_checkpoint(cfg, 3, "random")
    # the alternate spellings the vLLM weight mapper accepts (vllm/model_executor/models/dots_ocr.py hf_to_vllm_mapper)
    renamed = {k.replace(".attn.qkv.", ".attn.qkv_proj.").replace(".attn.proj.", ".attn.out_proj."): v for k, v in ckpt.items()}
    assert any(".attn.qkv_proj." in k for k in renamed)
    W.save_safetensors_dir(renamed, str(tmp_path), shards=3)
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".safetensors")]) == 3
    back = W.load_safetensors_dir(str(tmp_path))
    assert set(back) == set(ckpt)
    for k in ckpt:
        assert torch.equal(back[k], ckpt[k]), k
    W.validate_checkpoint(cfg, back)

    broken = dict(back)
    del broken["model.layers.1.mlp.down_proj.weight"]
    broken["lm_head.weight"] = broken["lm_head.weight"][:-1]
    with pytest.raises(ValueError, match=r"1 missing \[model\.layers\.1\.mlp\.down_proj\.weight\]; 1 mis-shaped \[lm_head\.weight"):
        W.validate_checkpoint(cfg, broken)
    with pytest.raises(ValueError, match="unexpected"):
        W.validate_checkpoint(cfg, dict(back, **{"model.rotary_emb.inv_freq": torch.zeros(4)}), allow_extra=False)
    W.validate_checkpoint(cfg, dict(back, **{"model.rotary_emb.inv_freq": torch.zeros(4)}))      # extras tolerated by default
    with pytest.raises(ValueError, match="mis-shaped"):
        W.validate_checkpoint(C.small(), back)
    (tmp_path / "empty").mkdir()
    with pytest.raises(FileNotFoundError):
        W.load_safetensors_dir(str(tmp_path / "empty"))


def test_hf_tokenizer_adapter_layout_stop_ids_and_decode(tmp_path):
    fast = _write_tokenizer(str(tmp_path))
    pad_id = fast.convert_tokens_to_ids("<|imgpad|>")
    eot, eoa = fast.convert_tokens_to_ids("<|endoftext|>"), fast.convert_tokens_to_ids("<|endofassistant|>")
    with open(tmp_path / "generation_config.json", "w") as f:
        json.dump({"eos_token_id": [eoa, eot], "pad_token_id": eot, "do_sample": False}, f)
    tk = HFTokenizer(str(tmp_path), image_token_id=pad_id)
    assert tk.stop_ids == (eoa, eot) and tk.eos_token_id == eoa and tk.pad_token_id == eot and tk.image_token_id == pad_id

    prompt = "Parse the page: ünïcode ✓"
    ids = tk.encode_chat(prompt, 7)
    whole = fast.encode("<|user|><|img|>" + "<|imgpad|>" * 7 + "<|endofimg|>" + prompt + "<|endofuser|><|assistant|>",
                        add_special_tokens=False)
    assert ids == whole                                   # piecewise tokenisation == tokenising the expanded string
    assert ids.count(pad_id) == 7
    assert tk.encode_chat(prompt, 1369).count(pad_id) == 1369      # second call comes from the per-prompt cache

    out = fast.encode("| a | b |\n", add_special_tokens=False)
    assert tk.decode(out + [eoa] + [eot] * 5) == "| a | b |\n"      # cut at the first stop id, pads dropped
    assert tk.decode(out + [eot, 65, 66]) == "| a | b |\n"
    assert tk.decode(out) == "| a | b |\n"

    batch = build_text_inputs(tk, [4, 9], [prompt, "x"])
    assert batch["input_ids"].shape == batch["attention_mask"].shape
    assert (batch["input_ids"] == pad_id).sum(dim=1).tolist() == [4, 9]
    assert batch["attention_mask"][1, 0].item() == 0 and batch["attention_mask"][:, -1].tolist() == [1, 1]   # left padded

    with pytest.raises(ValueError, match="image_token_id"):
        HFTokenizer(str(tmp_path), image_token_id=pad_id + 1)


def test_hf_tokenizer_uses_the_checkpoint_chat_template_when_present(tmp_path):
    template = ("{% for m in messages %}<|{{ m['role'] }}|>{{ m['content'] }}<|endof{{ m['role'] }}|>{% endfor %}"
                "{% if add_generation_prompt %}<|assistant|>{% endif %}")
    fast = _write_tokenizer(str(tmp_path), chat_template=template)
    tk = HFTokenizer(str(tmp_path))
    assert tk.render("hi") == "<|user|><|img|><|imgpad|><|endofimg|>hi<|endofuser|><|assistant|>"
    plain = HFTokenizer(_write_tokenizer(str(tmp_path / "plain")))
    assert plain.render("hi") == tk.render("hi")                   # the documented layout is the fallback
    assert plain.stop_ids == (fast.eos_token_id,)


def test_tokenizer_without_imgpad_is_rejected(tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {c: i for i, c in enumerate(sorted(pre_tokenizers.ByteLevel.alphabet()))}
    vocab["<unk>"] = len(vocab)
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[], unk_token="<unk>"))
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>")
    with pytest.raises(ValueError, match="imgpad"):
        HFTokenizer(fast)


class _RecordingEngine:
    """Stands in for Engine in from_checkpoint: records what it was built with and echoes a fixed completion."""

    def __init__(self, cfg, ckpt, device):
        self.cfg, self.ckpt, self.device = cfg, ckpt, torch.device("cpu")
        self.calls = []

    def generate(self, input_ids, attention_mask, max_new_tokens, eos_token_id, pad_token_id, **kw):
        self.calls.append(dict(eos=eos_token_id, pad=pad_token_id, n=max_new_tokens, keys=sorted(kw)))
        B = input_ids.shape[0]
        new = torch.full((B, max_new_tokens), pad_token_id, dtype=torch.int64)
        reply = self.reply
        new[:, :len(reply)] = torch.tensor(reply)
        new[:, len(reply)] = eos_token_id[0] if isinstance(eos_token_id, list) else eos_token_id
        from types import SimpleNamespace
        return SimpleNamespace(sequences=torch.cat([input_ids, new], dim=1))


def test_page_runner_from_checkpoint_directory(tmp_path):
    from PIL import Image
    from dots_ocr_b200.runner import PageRunner
    cfg = C.tiny()
    fast = _write_tokenizer(str(tmp_path))
    hf_cfg = _tiny_hf_config(cfg)
    hf_cfg["image_token_id"] = fast.convert_tokens_to_ids("<|imgpad|>")
    with open(tmp_path / "config.json", "w") as f:
        json.dump(hf_cfg, f)
    W.save_safetensors_dir(W.make_synthetic_checkpoint(cfg, 1, "random"), str(tmp_path), shards=2)

    runner = PageRunner.from_checkpoint(str(tmp_path), device="cpu", engine_factory=_RecordingEngine)
    eng = runner.engine
    assert eng.cfg.text == cfg.text and eng.cfg.image_token_id == hf_cfg["image_token_id"]
    assert set(eng.ckpt) == set(W.tensor_names(cfg))
    assert isinstance(runner.tokenizer, HFTokenizer)

    eng.reply = fast.encode('[{"bbox": [1, 2, 3, 4], "category": "Text", "text": "ok"}]', add_special_tokens=False)
    img = Image.new("RGB", (120, 90), (255, 255, 255))
    texts = runner.infer_batch([img, img], ["p1", "p2"], max_new_tokens=128, gpu_preprocess=False)
    assert texts == ['[{"bbox": [1, 2, 3, 4], "category": "Text", "text": "ok"}]'] * 2
    assert eng.calls[0]["eos"] == [fast.eos_token_id] and eng.calls[0]["n"] == 128
    assert "pixel_values" in eng.calls[0]["keys"] and "image_grid_thw" in eng.calls[0]["keys"]

    # a directory whose tensors do not fit its config.json is refused before any engine is built
    hf_cfg["num_hidden_layers"] = 3
    with open(tmp_path / "config.json", "w") as f:
        json.dump(hf_cfg, f)
    with pytest.raises(ValueError, match="missing"):
        PageRunner.from_checkpoint(str(tmp_path), device="cpu", engine_factory=_RecordingEngine)


def test_fabricated_checkpoint_directory_loads_everywhere(tmp_path):
    """tools/make_checkpoint_dir.py (tiny preset): our loader, the stock HF image processor and vLLM's DotsOCRConfig all read
    it; its tokenizer assigns the ids processing.SyntheticTokenizer uses, so both tokenizers drive the engine identically."""
    import importlib.util
    from dots_ocr_b200.processing import SyntheticTokenizer
    from dots_ocr_b200.runner import PageRunner
    spec = importlib.util.spec_from_file_location("make_checkpoint_dir",
                                                  os.path.join(os.path.dirname(__file__), "..", "tools", "make_checkpoint_dir.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    cfg = C.tiny()
    out = str(tmp_path / "DotsOCR")
    info = tool.write_dir(cfg, out, seed=5, flavour="peaked", shards=3)
    assert info["tensors"] == len(W.tensor_names(cfg))

    runner = PageRunner.from_checkpoint(out, device="cpu", engine_factory=_RecordingEngine)
    got = runner.engine.cfg
    assert (got.text, got.vision, got.image_token_id, got.video_token_id) == (cfg.text, cfg.vision, cfg.image_token_id,
                                                                              cfg.video_token_id)
    want = W.make_synthetic_checkpoint(cfg, 5, "peaked")
    assert all(torch.equal(runner.engine.ckpt[k], want[k]) for k in want)
    tk, st = runner.tokenizer, SyntheticTokenizer(cfg)
    prompt = "Please output: ünïcode ✓ 表格\n\t{}[]"
    assert tk.encode_chat(prompt, 9) == st.encode_chat(prompt, 9)
    assert tk.decode(list(prompt.encode())) == prompt == st.decode(list(prompt.encode()))
    sp = info["special_tokens"]
    assert tk.stop_ids == (sp["<|endoftext|>"], sp["<|endofassistant|>"]) and tk.pad_token_id == sp["<|endoftext|>"]
    with open(os.path.join(out, "model.safetensors.index.json")) as f:
        index = json.load(f)
    assert set(index["weight_map"]) == set(want) and set(index["weight_map"].values()) == \
        {f for f in os.listdir(out) if f.endswith(".safetensors")}

    from transformers import Qwen2VLImageProcessor
    ip = Qwen2VLImageProcessor.from_pretrained(out)
    assert (ip.patch_size, ip.merge_size, ip.size["shortest_edge"], ip.size["longest_edge"]) == (14, 2, 3136, 11289600)
    try:
        from vllm.transformers_utils.configs.dotsocr import DotsOCRConfig
    except Exception:
        return
    vc = DotsOCRConfig.from_pretrained(out)
    assert vc.vision_config.embed_dim == cfg.vision.embed_dim and vc.image_token_id == cfg.image_token_id
    assert vc.num_key_value_heads == cfg.text.num_key_value_heads and vc.architectures == ["DotsOCRForCausalLM"]


def test_build_inputs_equals_the_hf_processor_on_a_checkpoint_directory(tmp_path):
    """processing.build_inputs (HFTokenizer + our image path) against HF's Qwen2VLProcessor -- the class the checkpoint's
    DotsVLProcessor extends, configured the way vLLM configures it for dots.ocr (image token <|imgpad|>,
    vllm/model_executor/models/dots_ocr.py:149-160) -- on a fabricated directory: ids, mask, grid and pixels identical."""
    import importlib.util
    import numpy as np
    from PIL import Image
    from transformers import AutoTokenizer, Qwen2VLImageProcessor, Qwen2VLProcessor, Qwen2VLVideoProcessor
    from dots_ocr_b200.processing import build_inputs
    spec = importlib.util.spec_from_file_location("make_checkpoint_dir",
                                                  os.path.join(os.path.dirname(__file__), "..", "tools", "make_checkpoint_dir.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    cfg = C.tiny()
    out = str(tmp_path)
    with open(tmp_path / "config.json", "w") as f:
        json.dump(tool.hf_config_dict(cfg), f)
    tool.write_tokenizer(cfg, out)
    Qwen2VLImageProcessor(min_pixels=3136, max_pixels=11289600, patch_size=14, merge_size=2, temporal_patch_size=1).save_pretrained(out)

    tok = AutoTokenizer.from_pretrained(out)
    tok.image_token = "<|imgpad|>"
    tok.padding_side = "left"
    proc = Qwen2VLProcessor(image_processor=Qwen2VLImageProcessor.from_pretrained(out), tokenizer=tok,
                            video_processor=Qwen2VLVideoProcessor(), chat_template=tok.chat_template)
    proc.image_token, proc.video_token = "<|imgpad|>", "<|video_pad|>"
    tk = HFTokenizer(out, image_token_id=cfg.image_token_id)
    rng = np.random.default_rng(0)
    imgs = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in ((90, 130), (200, 120), (57, 400))]
    prompts = ["Parse the page", "Please output the layout: ünï 表", "x"]
    want = proc(text=[tk.render(p) for p in prompts], images=imgs, padding=True, return_tensors="pt")
    got = build_inputs(tk, imgs, prompts)
    for k in ("input_ids", "attention_mask", "image_grid_thw"):
        assert torch.equal(want[k], got[k].to(want[k].dtype)), k
    assert torch.equal(want["pixel_values"], got["pixel_values"])
    assert int((got["input_ids"] == cfg.image_token_id).sum()) == got["pixel_values"].shape[0] // 4
