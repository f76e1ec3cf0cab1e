"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU/PyTorch restatement of the reference's page-parsing arithmetic:

* ``oracle.vision``  -- the DotsVisionTransformer forward, restated line by line from the
  in-container mirror ``vllm/model_executor/models/dots_ocr.py:163-611`` (the reference
  repo itself contains no model code; its HF remote-code files are not available offline).
* ``oracle.model``   -- glue: HF ``Qwen2ForCausalLM`` (imported, not restated, from the
  installed ``transformers``) + ``GenerationMixin.generate(do_sample=False)``.

PARITY UNPINNED BY THE REFERENCE ITSELF: it ships no tests, golden tensors or token outputs
for this path, and neither its weights nor its HF-hub remote-code model
files exist offline.  What *is* pinned, all against code executed in this container:

* ``oracle.vision`` against vLLM 0.22's OWN ``DotsVisionTransformer`` (the implementation
  the reference's README tells users to serve with), run on CPU in fp32 on the same seeded
  weights: patch embed, every block and the merged embeddings agree to 2e-5 of range, at
  the tiny widths and at the real widths (1536 / 12 heads / 4224)
  (tests/golden/vllm_vision_tiny.npz, generator make_vllm_vision_golden.py);
* the decoder half is the reference's own dependency code (transformers' Qwen2ForCausalLM
  and generate) rather than a restatement;
* ``smart_resize`` and the post-decode functions against the reference functions themselves
  (tests/golden/smart_resize.json, postprocess.json).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.
"""
