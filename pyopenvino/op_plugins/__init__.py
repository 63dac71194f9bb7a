"""Top-level `op_plugins` package for `sys.path.append('pyopenvino'); import op_plugins.Convolution as op`
(`test_node_sample.py:3,11`): every `op_plugins.<Type>` is the module `pyopenvino_b200.op_plugins.<Type>`."""
import importlib
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

_impl = importlib.import_module('pyopenvino_b200.op_plugins')
for _f in sorted(os.listdir(os.path.dirname(_impl.__file__))):
    if _f.endswith('.py') and not _f.startswith('_'):
        _m = importlib.import_module('pyopenvino_b200.op_plugins.' + _f[:-3])
        sys.modules[__name__ + '.' + _f[:-3]] = _m
        globals()[_f[:-3]] = _m
