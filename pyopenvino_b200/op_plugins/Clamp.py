"""Clamp plugin -- drop-in for `op_plugins/Clamp.py` (`np.clip(x, float(min), float(max))`, `Clamp.py:9-12,45-46`)."""
from .. import common_def, kernels, plugin_util


def name():
    print('Clamp')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    max_val = float(node['data']['max'])
    min_val = float(node['data']['min'])
    f = fused or {}
    y = kernels.affine_act(inputs[0], act=('clamp', min_val, max_val), out=f.get('out'))
    return plugin_util.finish(node, inputs, y)
