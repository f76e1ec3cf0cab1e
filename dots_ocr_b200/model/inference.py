"""Page -> text through the in-process B200 engine, behind the reference's model-adapter seam.

``inference_with_vllm`` keeps the reference signature (``dots_ocr/model/inference.py:7-18``) so that
``from dots_ocr_b200.model.inference import inference_with_vllm`` is a drop-in for callers such as
``DotsOCRParser._inference_with_vllm`` (``dots_ocr/parser.py:119-131``) and ``demo/demo_vllm.py:35``.
The HTTP transport arguments are accepted and ignored: the request is served by the engine living in
this process (one per GPU).  Like the reference, a transport-level failure returns ``None``; engine
errors propagate as Python exceptions (the reference's OpenAI client raises too).
"""
from __future__ import annotations

import threading
from typing import Optional

_state = {"runner": None}
_lock = threading.Lock()


def set_default_runner(runner) -> None:
    """Install the process-wide page runner (see ``dots_ocr_b200.runner.PageRunner``)."""
    with _lock:
        _state["runner"] = runner


def get_default_runner():
    """Process-wide runner: the engine behind a request batcher, so that the parser's thread fan-out (one page per call,
    parser.py:282-290) shares generate() calls the way requests share a vLLM server's batches."""
    with _lock:
        if _state["runner"] is None:
            from ..batching import BatchingRunner
            from ..runner import PageRunner
            _state["runner"] = BatchingRunner(PageRunner.from_default())
        return _state["runner"]


def inference_with_vllm(
        image,
        prompt,
        protocol="http",
        ip="localhost",
        port=8000,
        temperature=0.1,
        top_p=0.9,
        max_completion_tokens=32768,
        model_name="rednote-hilab/dots.mocr",
        system_prompt=None,
        ) -> Optional[str]:
    """Greedy page inference.  ``temperature``/``top_p`` are accepted for signature compatibility; the
    engine decodes greedily (BASELINE.json fixes greedy), which is the temperature -> 0 limit."""
    runner = get_default_runner()
    text = prompt if system_prompt is None else f"{system_prompt}\n{prompt}"
    return runner.infer(image, text, max_new_tokens=max_completion_tokens)


# explicit alias under the engine's own name
inference_with_b200 = inference_with_vllm
