"""`pyopenvino` import alias: scripts written for the reference run unchanged against the B200 engine.

    from pyopenvino.inference_engine import IECore          # test_pyopenvino.py:7, test_googlenet_v1.py, integrity_test.py
    sys.path.append('pyopenvino'); import common_def        # pyopenvino/inference_engine.py:17-18
    sys.path.append('pyopenvino'); import op_plugins.Convolution as op   # test_node_sample.py:3,11

Nothing lives here: every name is the module of the same name in `pyopenvino_b200` (one module object, two names).
"""
import importlib
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import pyopenvino_b200 as _impl  # noqa: E402

__version__ = _impl.__version__


def _alias(alias, target):
    mod = importlib.import_module(target)
    sys.modules[alias] = mod
    return mod


common_def = _alias(__name__ + '.common_def', 'pyopenvino_b200.common_def')
inference_engine = _alias(__name__ + '.inference_engine', 'pyopenvino_b200.inference_engine')
op_plugins = _alias(__name__ + '.op_plugins', 'pyopenvino_b200.op_plugins')
for _f in sorted(os.listdir(os.path.dirname(op_plugins.__file__))):
    if _f.endswith('.py') and not _f.startswith('_'):
        _alias(__name__ + '.op_plugins.' + _f[:-3], 'pyopenvino_b200.op_plugins.' + _f[:-3])
