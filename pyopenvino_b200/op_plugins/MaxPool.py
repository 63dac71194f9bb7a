"""MaxPool plugin -- drop-in for `op_plugins/MaxPool.py`.

Semantics of the reference 'numpy' kernel (`MaxPool.py:41-72`): zero padding takes part in the max,
ceil-mode windows overhanging the padded tensor are clipped, `same_*` keeps the input size
(`MaxPool.py:34-36`).  A per-channel Multiply + Add that follows (folded BatchNorm in mnist_bn) can
be folded into the kernel through `fused`.
"""
from .. import _cabi, common_def, kernels, plugin_util


def name():
    print('MaxPool')


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
    y = kernels.pool2d(x, _cabi.POOL_MAX, kernel, strides, pads_begin, pads_end, out_hw, scale=f.get('scale'),
                       shift=f.get('shift'), out=f.get('out'))
    return plugin_util.finish(node, inputs, y)
