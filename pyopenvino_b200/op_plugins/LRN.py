"""LRN plugin -- drop-in for `op_plugins/LRN.py` (across-channel, `LRN.py:10-22`).

Matches the reference, not the OpenVINO spec: alpha is not divided by `size` and the axes input
(port 1) is ignored.
"""
from .. import common_def, kernels, plugin_util


def name():
    print('LRN')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    data = node['data']
    f = fused or {}
    y = kernels.lrn(inputs[0], int(data['size']), float(data['alpha']), float(data['beta']), float(data['bias']),
                    out=f.get('out'))
    return plugin_util.finish(node, {0: inputs[0]}, y)
