"""Result plugin -- drop-in for `op_plugins/Result.py`: the device -> host edge of the graph.

Stores the host ndarray (logical layout) in `node['result']` and returns `[]` like the reference
(`Result.py:17-18`).
"""
import numpy as np

from .. import common_def
from ..device import is_device


def name():
    print('Result')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    x = inputs[0]
    node['result'] = x.numpy() if is_device(x) else np.asarray(x)
    return []
