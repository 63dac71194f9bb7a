"""Concat plugin -- drop-in for `op_plugins/Concat.py` (`np.concatenate` along `axis` over
`inputs.values()`, i.e. in the order the executor filled the dict = edge order of the IR, which is port order in
every shipped model but is not sorted by port -- exactly like the reference, `Concat.py:9-13`).

On the device a channel concat of NHWC feature maps is a set of strided row copies
(`b200ov_copy2d`) -- or nothing at all when the executor already made the producers write into
channel slices of the output buffer (`fused['inplace']`).  Host int64 / constant inputs (the SSD
prior-box branch, folded once at load time) are concatenated on the host.
"""
import ctypes as C

import numpy as np

from .. import _cabi, common_def, kernels, plugin_util
from .. import device as dev
from ..device import DeviceArray, is_device


def name():
    print('Concat')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    axis = int(node['data']['axis'])
    assert len(inputs) > 1
    assert axis <= inputs[0].ndim
    f = fused or {}
    if f.get('inplace') is not None:
        return {common_def.first_output_port(node): f['inplace']}      # producers already wrote their slices
    parts = list(inputs.values())
    if not any(is_device(p) for p in parts) and parts[0].dtype != np.float32:
        return {common_def.first_output_port(node): np.concatenate(parts, axis=axis)}   # shape arithmetic on the host
    parts = [kernels.as_device(p) for p in parts]
    shape = list(parts[0].shape)
    shape[axis] = sum(p.shape[axis] for p in parts)
    if len(shape) == 4 and axis == 1 and all(p.layout == 'nhwc' for p in parts):
        n, c, h, w = shape
        out = kernels.new_nhwc(n, c, h, w)
        off = 0
        for p in parts:
            kernels.copy_channels(p, kernels.channel_slice(out, off, p.shape[1]))
            off += p.shape[1]
    else:
        parts = [kernels.as_plain(p) for p in parts]
        outer = int(np.prod(shape[:axis])) if axis > 0 else 1
        inner_total = int(np.prod(shape[axis:]))
        out = DeviceArray(dev.alloc_f32(outer * inner_total), shape, 'plain')
        inners = [int(np.prod(p.shape[axis:])) for p in parts]
        if len(parts) <= _cabi.CONCAT_MAX_PARTS:
            # one launch for all parts (the SSD head Concats have six)
            srcs = (C.c_void_p * len(parts))(*[p.ptr for p in parts])
            cols = (C.c_int * len(parts))(*inners)
            _cabi.call('b200ov_concat_rows', len(parts), srcs, cols, C.c_void_p(out.ptr), outer, C.c_void_p(dev.stream()))
        else:
            off = 0
            for p, inner in zip(parts, inners):
                _cabi.call('b200ov_copy2d', C.c_void_p(p.ptr), C.c_void_p(out.ptr + 4 * off), outer, inner, inner, inner_total,
                           C.c_void_p(dev.stream()))
                off += inner
    return plugin_util.finish(node, inputs, out)
