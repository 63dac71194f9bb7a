"""StridedSlice plugin -- drop-in for `op_plugins/StridedSlice.py`: `x[b:e:s, ...]` with every mask
ignored (`StridedSlice.py:8-24`).  Only used on host shape vectors in the SSD prior-box branch."""
import numpy as np

from .. import _cabi, common_def
from ..device import is_device


def name():
    print('StridedSlice')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    x = inputs[0]
    if is_device(x):
        raise _cabi.B200ovError('StridedSlice on a device tensor has no kernel (only host shape vectors are sliced)')
    begin, end, stride = (np.asarray(inputs[i]).reshape(-1) for i in (1, 2, 3))
    idx = tuple(slice(int(begin[d]), int(end[d]), int(stride[d])) for d in range(np.asarray(x).ndim))
    return {common_def.first_output_port(node): np.asarray(x)[idx]}
