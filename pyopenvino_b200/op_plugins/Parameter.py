"""Parameter plugin -- drop-in for `op_plugins/Parameter.py`: the host -> device edge of the graph.

Like the reference (`Parameter.py:11-13`) the user array is reshaped to the IR shape and cast to the
element type; it is then copied to HBM.  A scalar / per-channel Multiply and Add that follow the
input (GoogLeNet `data/mean`, SSD `Preprocessor/mul` + `/sub`) can be folded into the NCHW -> NHWC
conversion kernel through `fused`.
"""
import numpy as np

from .. import common_def, kernels
from ..device import is_device


def name():
    print('Parameter')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    shape = node['data']['shape']
    precision = common_def.type_convert_tbl[node['data']['element_type']]
    param = node['param']
    if is_device(param):                      # executor staged the batch in a static device buffer already
        assert tuple(param.shape) == tuple(shape)
        x = param
    else:
        x = kernels.upload(np.array(param).reshape(shape).astype(precision))
    f = fused or {}
    if x.ndim == 4 and (f.get('scale') is not None or f.get('shift') is not None or f.get('to_nhwc')):
        x = kernels.to_nhwc(x, scale=f.get('scale'), shift=f.get('shift'))
    return {0: x}
