from .prompts import dict_promptmode_to_prompt  # noqa: F401  (reference: dots_ocr/utils/__init__.py:1)
from .layout_utils import post_process_output, post_process_cells, pre_process_bboxes  # noqa: F401,E402
from .format_transformer import layoutjson2md  # noqa: F401,E402
