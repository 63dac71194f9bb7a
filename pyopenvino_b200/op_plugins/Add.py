"""Add plugin -- drop-in for `op_plugins/Add.py` (`in0 + broadcast_to(in1, in0.shape)`, `Add.py:9-14`).

Standalone kernel for nodes the executor could not fold into a producer epilogue: a scalar or
per-channel operand goes through `b200ov_affine_act`, two same-shape tensors through `b200ov_binary`.
"""
import numpy as np

from .. import _cabi, common_def, kernels, plugin_util


def name():
    print('Add')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    a, b = inputs[0], inputs[1]
    # port 1 must broadcast to port 0's shape (np.broadcast_to raises otherwise, Add.py:12)
    if np.broadcast_shapes(tuple(b.shape), tuple(a.shape)) != tuple(a.shape):
        raise ValueError('Add: operand of shape {} cannot be broadcast to {}'.format(tuple(b.shape), tuple(a.shape)))
    f = fused or {}
    a = kernels.as_device(a)
    if a.ndim == 4 and a.layout == 'plain' and b.size > 1 and tuple(b.shape) != tuple(a.shape):
        a = kernels.to_nhwc(a)        # per-channel operand: work on the NHWC feature map
    if tuple(a.shape) == tuple(b.shape) and b.size > 1:
        y = kernels.binary(0, a, b)
        if f.get('act') is not None:
            y = kernels.affine_act(y, act=f['act'])
    elif kernels._channel_operand_ok(a, b):
        y = kernels.affine_act(a, shift=b, act=f.get('act'), out=f.get('out'))
    else:
        raise _cabi.B200ovError('Add: broadcast {} -> {} has no device kernel'.format(tuple(b.shape), tuple(a.shape)))
    return plugin_util.finish(node, inputs, y)
