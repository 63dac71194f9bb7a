"""Transpose plugin -- drop-in for `op_plugins/Transpose.py` (`x.transpose(axes)`, `Transpose.py:9-13`).

Feature maps already live in HBM as NHWC, so the NCHW -> NHWC permutation every model uses
(`[0,2,3,1]`: before the flatten in the MNIST nets, on the 12 SSD head outputs) is a change of
metadata; other permutations run the tiled transpose kernel (`b200ov_transpose`).
"""
import ctypes as C

import numpy as np

from .. import _cabi, common_def, kernels, plugin_util
from .. import device as dev
from ..device import DeviceArray


def name():
    print('Transpose')


def _transpose_batched(x, batch, rows, cols, out_shape):
    out = DeviceArray(dev.alloc_f32(batch * rows * cols), out_shape, 'plain')
    _cabi.call('b200ov_transpose', C.c_void_p(x.ptr), C.c_void_p(out.ptr), batch, rows, cols, cols, rows,
               C.c_void_p(dev.stream()))
    return out


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    perm = [int(p) for p in np.asarray(inputs[1]).reshape(-1)]
    x = kernels.as_device(inputs[0])
    assert sorted(perm) == list(range(x.ndim)), 'bad permutation {}'.format(perm)
    out_shape = tuple(x.shape[p] for p in perm)
    if perm == list(range(x.ndim)):
        y = kernels.as_plain(x)
    elif x.ndim == 4 and perm == [0, 2, 3, 1]:
        if x.layout == 'nhwc':
            if x.is_dense() and x.st == 'f32':
                y = DeviceArray(x.t, out_shape, 'plain')          # zero-copy: physical layout already is N,H,W,C
            else:                                                 # channel slice, or FP16 storage: one (widening) copy
                n, c, h, w = x.shape
                y = DeviceArray(dev.alloc_f32(x.size), out_shape, 'plain')
                _cabi.call('b200ov_copy2d_st', C.c_void_p(x.ptr), x.code, C.c_void_p(y.ptr), _cabi.DT_F32, n * h * w, c, x.ld, c,
                           C.c_void_p(dev.stream()))
        else:
            n, c, h, w = x.shape
            y = _transpose_batched(x, n, c, h * w, out_shape)
    elif x.ndim == 4 and perm == [0, 3, 1, 2]:
        x = kernels.as_plain(x)
        n, h, w, c = x.shape
        y = _transpose_batched(x, n, h * w, c, out_shape)
    elif x.ndim == 2 and perm == [1, 0]:
        x = kernels.as_plain(x)
        y = _transpose_batched(x, 1, x.shape[0], x.shape[1], out_shape)
    else:
        raise _cabi.B200ovError('Transpose: permutation {} of a {}-D tensor has no device kernel'.format(perm, x.ndim))
    return plugin_util.finish(node, {0: inputs[0]}, y)
