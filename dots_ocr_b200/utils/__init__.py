from .prompts import dict_promptmode_to_prompt  # noqa: F401  (reference: dots_ocr/utils/__init__.py:1)
