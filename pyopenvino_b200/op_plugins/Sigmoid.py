"""Sigmoid plugin -- drop-in for `op_plugins/Sigmoid.py` (standalone elementwise kernel, `b200ov_affine_act`)."""
from .. import common_def, kernels, plugin_util


def name():
    print('Sigmoid')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    f = fused or {}
    y = kernels.affine_act(inputs[0], act=('sigmoid',), out=f.get('out'))
    return plugin_util.finish(node, inputs, y)
