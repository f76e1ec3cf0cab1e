"""Task prompts (input data of the hot path: the prompt text fixes the prefill length).

The strings are the reference's (``dots_ocr/utils/prompts.py:1-46``), kept as a
data file so they stay byte-identical.
"""
import json
import os

with open(os.path.join(os.path.dirname(__file__), "prompts.json"), encoding="utf-8") as _f:
    dict_promptmode_to_prompt = json.load(_f)
