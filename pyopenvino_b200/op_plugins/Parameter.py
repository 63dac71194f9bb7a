"""Parameter plugin -- drop-in for `op_plugins/Parameter.py`: the host -> device edge of the graph.

Like the reference (`Parameter.py:11-13`) the user array is reshaped to the IR shape and cast to the
element type; it is then copied to HBM.  A scalar / per-channel Multiply and Add that follow the
input (GoogLeNet `data/mean`, SSD `Preprocessor/mul` + `/sub`) can be folded into the NCHW -> NHWC
conversion kernel through `fused`.
"""
import numpy as np

from .. import _cabi, common_def, kernels
from ..device import DeviceArray, RawInput, is_device, native_input


def name():
    print('Parameter')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    shape = node['data']['shape']
    precision = common_def.type_convert_tbl[node['data']['element_type']]
    param = node['param']
    f = fused or {}
    to_nhwc = f.get('scale') is not None or f.get('shift') is not None or f.get('to_nhwc')
    if is_device(param):                      # executor staged the batch in a static device buffer already
        assert tuple(param.shape) == tuple(shape)
        x = param
    elif isinstance(param, RawInput) or precision is np.float32:
        # the input crosses PCIe in its native width (uint8 / int8 / float16 / float32); the cast to the IR precision
        # (`.astype(precision)`, Parameter.py:13) happens in the layout kernel and is exact for these types
        raw = param if isinstance(param, RawInput) else kernels.upload_raw(np.asarray(native_input(param)).reshape(shape))
        assert tuple(raw.shape) == tuple(shape)
        if raw.np_dtype == np.float32 and not (raw.ndim == 4 and to_nhwc):
            return {0: DeviceArray(raw.t, raw.shape, 'plain')}
        return {0: kernels.input_to_device(raw, scale=f.get('scale'), shift=f.get('shift'), nhwc=bool(to_nhwc),
                                           split=bool(f.get('split')) and kernels.default_math == _cabi.MATH_AUTO and
                                           kernel_type not in ('fp32', 'tf32x3', 'tf32', 'safe'))}
    else:
        x = kernels.upload(np.array(param).reshape(shape).astype(precision))
    if x.ndim == 4 and to_nhwc:
        x = kernels.to_nhwc(x, scale=f.get('scale'), shift=f.get('shift'))
    return {0: x}
