"""Top-level `common_def` for callers that put `./pyopenvino` on sys.path like the reference does
(`pyopenvino/inference_engine.py:17-18`, every `op_plugins/<Type>.py`): the names of
`pyopenvino_b200.common_def`."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from pyopenvino_b200.common_def import *  # noqa: E402,F401,F403
from pyopenvino_b200.common_def import format_config, type_convert_tbl  # noqa: E402,F401
