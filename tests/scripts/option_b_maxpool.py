# INTEGRATION.md option B, verbatim: what a maintainer of the reference would add to pyopenvino/op_plugins/MaxPool.py to
# run its arithmetic on libb200ov through raw ctypes.  Nothing from pyopenvino_b200's Python side is imported.
import ctypes as C, numpy as np
import math, os, sys

_lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'pyopenvino_b200', 'libb200ov.so'))
_lib.b200ov_last_error.restype = C.c_char_p

class PoolDesc(C.Structure):                    # mirrors b200ov_pool_desc in include/b200ov.h
    _fields_ = [(n, C.c_int32) for n in ('n','h','w','c','kh','kw','sh','sw','pt','pl','pb','pr','oh','ow','x_ld','y_ld','mode','dtype')]

def _check(rc):
    if rc != 0:
        raise RuntimeError(_lib.b200ov_last_error().decode())

def calc_output_shape(input_dim, kernel_dim, strides, pads_begin, pads_end, rounding_type, auto_pad):
    # the maintainer already has this function (MaxPool.py:10-38); restated for a stand-alone script
    rnd = math.floor if rounding_type == 'floor' else math.ceil
    if auto_pad == 'explicit':
        return tuple(rnd((i + pb + pe - k) / s) + 1 for i, k, s, pb, pe in zip(input_dim, kernel_dim, strides, pads_begin, pads_end))
    if auto_pad == 'valid':
        return tuple(rnd((i - k) / s) + 1 for i, k, s in zip(input_dim, kernel_dim, strides))
    return tuple(input_dim)

def kernel_MaxPool_b200(inputs, strides, pads_begin, pads_end, kernel, rounding_type, auto_pad):
    x = np.ascontiguousarray(inputs[0].transpose(0, 2, 3, 1))            # NCHW -> NHWC on the host (or b200ov_transpose)
    n, h, w, c = x.shape
    oh, ow = calc_output_shape((h, w), kernel, strides, pads_begin, pads_end, rounding_type, auto_pad)   # MaxPool.py:10-38
    y = np.empty((n, oh, ow, c), np.float32)
    dx, dy = C.c_void_p(), C.c_void_p()
    _check(_lib.b200ov_init(0))
    _check(_lib.b200ov_malloc(C.byref(dx), C.c_size_t(x.nbytes)))
    _check(_lib.b200ov_malloc(C.byref(dy), C.c_size_t(y.nbytes)))
    _check(_lib.b200ov_memcpy_h2d(dx, x.ctypes.data_as(C.c_void_p), C.c_size_t(x.nbytes), None))
    d = PoolDesc(n, h, w, c, kernel[0], kernel[1], strides[0], strides[1], pads_begin[0], pads_begin[1],
                 pads_end[0], pads_end[1], oh, ow, c, c, 0, 0)            # mode 0 = B200OV_POOL_MAX, dtype 0 = float32
    _check(_lib.b200ov_pool2d(C.byref(d), dx, None, None, dy, None))
    _check(_lib.b200ov_memcpy_d2h(y.ctypes.data_as(C.c_void_p), dy, C.c_size_t(y.nbytes), None))
    _check(_lib.b200ov_stream_sync(None))
    _lib.b200ov_free(dx); _lib.b200ov_free(dy)
    return y.transpose(0, 3, 1, 2)                                       # back to the reference's NCHW

if __name__ == '__main__':
    assert 'pyopenvino_b200' not in sys.modules and 'torch' not in sys.modules
    x = np.load(sys.argv[1])
    args = [tuple(int(v) for v in a.split(',')) for a in sys.argv[3:7]]
    y = kernel_MaxPool_b200({0: x}, args[0], args[1], args[2], args[3], sys.argv[7], sys.argv[8])
    assert 'torch' not in sys.modules
    np.save(sys.argv[2], np.ascontiguousarray(y))
