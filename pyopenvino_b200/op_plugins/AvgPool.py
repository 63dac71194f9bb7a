"""AvgPool plugin -- drop-in for `op_plugins/AvgPool.py`.

Reproduces the reference 'numpy' kernel (`AvgPool.py:41-59`) including its quirk: pads and
`exclude-pad` are ignored and the window is clipped at h-1 / w-1 (`AvgPool.py:56`), so GoogLeNet's
7x7 pool averages a 6x6 window.
"""
from .. import _cabi, common_def, kernels, plugin_util


def name():
    print('AvgPool')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    data = node['data']
    strides = common_def.string_to_tuple(data['strides'])
    pads_begin = common_def.string_to_tuple(data['pads_begin'])
    pads_end = common_def.string_to_tuple(data['pads_end'])
    kernel = common_def.string_to_tuple(data['kernel'])
    x = inputs[0]
    n, c, h, w = x.shape
    out_hw = common_def.spatial_output_shape((h, w), kernel, strides, pads_begin, pads_end, data['rounding_type'],
                                             data['auto_pad'], False)
    f = fused or {}
    y = kernels.pool2d(x, _cabi.POOL_AVG_REF, kernel, strides, pads_begin, pads_end, out_hw, out=f.get('out'))
    return plugin_util.finish(node, inputs, y)
