"""Unsqueeze plugin -- drop-in for `op_plugins/Unsqueeze.py` (`np.expand_dims`, `Unsqueeze.py:9-14`); metadata only."""
import numpy as np

from .. import common_def, kernels, plugin_util
from ..device import DeviceArray, is_device


def name():
    print('Unsqueeze')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    x = inputs[0]
    axes = [int(a) for a in np.asarray(inputs[1]).reshape(-1)]
    if not is_device(x):
        return {common_def.first_output_port(node): np.expand_dims(np.asarray(x), axes)}
    x = kernels.as_plain(x)
    shape = np.expand_dims(np.empty(x.shape, dtype=np.int8), axes).shape
    return {common_def.first_output_port(node): DeviceArray(x.t, shape, 'plain')}
