"""Multiply plugin -- drop-in for `op_plugins/Multiply.py`.

The operand with fewer elements is broadcast to the other one's shape (`Multiply.py:9-17`); as in
the reference (`Multiply.py:46-62`) `kernel_type` does not change the result.
"""
import numpy as np

from .. import _cabi, common_def, kernels, plugin_util


def name():
    print('Multiply')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    a, b = inputs[0], inputs[1]
    if not a.size > b.size:
        a, b = b, a          # broadcast port 0 to port 1's shape (Multiply.py:14-15); product commutes
    if np.broadcast_shapes(tuple(b.shape), tuple(a.shape)) != tuple(a.shape):
        raise ValueError('Multiply: operand of shape {} cannot be broadcast to {}'.format(tuple(b.shape), tuple(a.shape)))
    f = fused or {}
    a = kernels.as_device(a)
    if a.ndim == 4 and a.layout == 'plain' and b.size > 1 and tuple(b.shape) != tuple(a.shape):
        a = kernels.to_nhwc(a)        # per-channel operand: work on the NHWC feature map
    if tuple(a.shape) == tuple(b.shape) and b.size > 1:
        y = kernels.binary(1, a, b)
        if f.get('act') is not None:
            y = kernels.affine_act(y, act=f['act'])
    elif kernels._channel_operand_ok(a, b):
        y = kernels.affine_act(a, scale=b, shift=f.get('shift'), act=f.get('act'),
                               out=f.get('out'))
    else:
        raise _cabi.B200ovError('Multiply: broadcast {} -> {} has no device kernel'.format(tuple(b.shape), tuple(a.shape)))
    return plugin_util.finish(node, inputs, y)
