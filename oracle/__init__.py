"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU/PyTorch restatement of the reference's page-parsing arithmetic:

* ``oracle.vision``  -- the DotsVisionTransformer forward, restated line by line from the
  in-container mirror ``vllm/model_executor/models/dots_ocr.py:163-611`` (the reference
  repo itself contains no model code; its HF remote-code files are not available offline).
* ``oracle.model``   -- glue: HF ``Qwen2ForCausalLM`` (imported, not restated, from the
  installed ``transformers``) + ``GenerationMixin.generate(do_sample=False)``.

PARITY UNPINNED: the reference ships no tests, golden tensors or token outputs for this
path, and neither its weights nor its remote-code model files exist
offline.  What *is* pinned: ``smart_resize`` against the reference function executed in
this container (tests/golden/smart_resize.json), and the decoder half, which is the
reference's own dependency code (transformers' Qwen2) rather than a restatement.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.
"""
