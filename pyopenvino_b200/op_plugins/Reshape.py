"""Reshape plugin -- drop-in for `op_plugins/Reshape.py`.

Target-shape rules of the reference (`Reshape.py:14-44`): 0 copies the input dimension (only while
left-aligned), one -1 is inferred, `special_zero` is ignored.  On the device this is metadata only;
an NHWC feature map with H*W > 1 is first brought to the logical NCHW order (`b200ov_transpose`).
"""
import numpy as np

from .. import common_def, kernels, plugin_util
from ..device import DeviceArray, is_device


def name():
    print('Reshape')


def resolve_shape(in_shape, target):
    size = int(np.prod(in_shape)) if len(in_shape) else 1
    dims, deferred, zero_ok = [], -1, True
    for idx, d in enumerate(int(t) for t in target):
        if d == 0:
            assert zero_ok, "'0' must be left aligned in a Reshape target"
            d0 = int(in_shape[idx])
            assert size % d0 == 0
            dims.append(d0)
            size //= d0
        else:
            zero_ok = False
            if d == -1:
                assert deferred == -1, 'more than one -1 in a Reshape target'
                deferred = idx
                dims.append(-1)
            else:
                assert size % d == 0
                dims.append(d)
                size //= d
    if deferred != -1:
        dims[deferred] = int(size)
    elif size != 1:
        # the reference's `input0.reshape(adjusted_dims)` raises here (`Reshape.py:44`): a target without -1 must keep
        # the element count (e.g. a hard-coded batch-1 target on a re-batched network)
        raise ValueError('cannot reshape array of shape {} into shape {}'.format(tuple(in_shape), tuple(dims)))
    return tuple(dims)


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    x = inputs[0]
    shape = resolve_shape(tuple(x.shape), np.asarray(inputs[1]).reshape(-1))
    if not is_device(x) and x.dtype != np.float32:
        return {common_def.first_output_port(node): np.asarray(x).reshape(shape)}
    x = kernels.as_plain(x)
    y = DeviceArray(x.t, shape, 'plain')
    return plugin_util.finish(node, {0: inputs[0]}, y)
