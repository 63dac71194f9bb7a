"""ctypes binding of libb200ov.so (the C ABI declared in include/b200ov.h).

There is no CPU fallback: importing this module without the built library, or calling a kernel
without a CUDA device, raises.  Build with `python -m pyopenvino_b200.build`.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libb200ov.so')

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_CLAMP, ACT_SIGMOID = 0, 1, 2, 3
MATH_AUTO, MATH_FP32, MATH_TF32X3, MATH_TF32, MATH_F16X2, MATH_SAFE = 0, 1, 2, 3, 4, 5
POOL_MAX, POOL_AVG_REF = 0, 1
DT_F32, DT_F16, DT_U8, DT_I8, DT_HL = 0, 1, 2, 3, 4
CONCAT_MAX_PARTS = 8                 # B200OV_CONCAT_MAX_PARTS


class B200ovError(RuntimeError):
    code = None               # the library's status code when the error came from an entry point


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n', 'h', 'w', 'cin', 'cout', 'kh', 'kw', 'sh', 'sw', 'pt', 'pl', 'oh', 'ow',
                                         'x_ld', 'y_ld', 'ldw', 'act')] + \
               [('act_lo', C.c_float), ('act_hi', C.c_float), ('math', C.c_int32), ('x_dtype', C.c_int32), ('y_dtype', C.c_int32),
                ('pre_pool', C.c_int32)]


PREPOOL_NONE, PREPOOL_MAX3X3S1 = 0, 1


class DwConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n', 'h', 'w', 'c', 'kh', 'kw', 'sh', 'sw', 'pt', 'pl', 'oh', 'ow',
                                         'x_ld', 'y_ld', 'act')] + [('act_lo', C.c_float), ('act_hi', C.c_float),
                                                                    ('math', C.c_int32), ('dtype', C.c_int32),
                                                                    ('y_dtype', C.c_int32)]


DW_AUTO, DW_EXACT = 0, 1


class ConvSeg(C.Structure):
    _fields_ = [('y', C.c_void_p), ('col0', C.c_int32), ('cout', C.c_int32), ('y_ld', C.c_int32), ('y_dtype', C.c_int32)]


class PoolDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n', 'h', 'w', 'c', 'kh', 'kw', 'sh', 'sw', 'pt', 'pl', 'pb', 'pr', 'oh', 'ow',
                                         'x_ld', 'y_ld', 'mode', 'dtype')]


class DetectionDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('n', 'num_priors', 'num_classes', 'keep_top_k', 'code_center_size',
                                         'variance_encoded_in_target', 'clip_before_nms', 'clip_after_nms')] + \
               [('confidence_threshold', C.c_float), ('nms_threshold', C.c_float)]


_P, _I, _F, _L, _Z = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t

# name -> argtypes; every function returns int except the two noted below
SIGNATURES = {
    'b200ov_init': [_I],
    'b200ov_device_info': [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_Z)],
    'b200ov_malloc': [C.POINTER(_P), _Z],
    'b200ov_free': [_P],
    'b200ov_host_alloc': [C.POINTER(_P), _Z],
    'b200ov_host_free': [_P],
    'b200ov_memcpy_h2d': [_P, _P, _Z, _P],
    'b200ov_memcpy_d2h': [_P, _P, _Z, _P],
    'b200ov_memcpy_d2d': [_P, _P, _Z, _P],
    'b200ov_memset': [_P, _I, _Z, _P],
    'b200ov_stream_create': [C.POINTER(_P)],
    'b200ov_stream_destroy': [_P],
    'b200ov_stream_sync': [_P],
    'b200ov_graph_begin': [_P],
    'b200ov_graph_end': [_P, C.POINTER(_P)],
    'b200ov_graph_launch': [_P, _P],
    'b200ov_graph_destroy': [_P],
    'b200ov_conv_weight_dims': [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_L)],
    'b200ov_pack_conv_weights': [_P, _P, _I, _I, _I, _I, _P],
    'b200ov_conv2d': [C.POINTER(ConvDesc), _P, _P, _P, _P, _P],
    'b200ov_conv2d_multi': [C.POINTER(ConvDesc), _P, _P, _P, _I, C.POINTER(ConvSeg), _P],
    'b200ov_matmul': [_I, _I, _I, _P, _I, _P, _I, _P, _I, _F, _F, _I, _P, _I, _P],
    'b200ov_matmul_workspace': [_I, _I, _I, C.POINTER(_Z)],
    'b200ov_matmul_ws': [_I, _I, _I, _P, _I, _P, _I, _P, _I, _F, _F, _I, _P, _I, _P, _Z, _P],
    'b200ov_status_word': [C.POINTER(C.c_void_p)],
    'b200ov_status_reset': [_P],
    'b200ov_status_fetch': [_P, _P],
    'b200ov_pack_dw_weights': [_P, _P, _I, _I, _I, _P],
    'b200ov_dwconv2d': [C.POINTER(DwConvDesc), _P, _P, _P, _P, _P],
    'b200ov_pool2d': [C.POINTER(PoolDesc), _P, _P, _P, _P, _P],
    'b200ov_affine_act': [_P, _P, _L, _I, _I, _I, _I, _P, _F, _I, _P, _F, _I, _F, _F, _P],
    'b200ov_binary': [_I, _P, _P, _P, _L, _P],
    'b200ov_softmax': [_P, _P, _I, _I, _P],
    'b200ov_lrn': [_P, _P, _L, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200ov_lrn_st': [_P, _P, _I, _L, _I, _I, _I, _I, _F, _F, _F, _P],
    'b200ov_transpose': [_P, _P, _I, _I, _I, _I, _I, _P],
    'b200ov_transpose_st': [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    'b200ov_copy2d_st': [_P, _I, _P, _I, _L, _I, _I, _I, _P],
    'b200ov_nchw_to_nhwc_affine': [_P, _P, _I, _I, _I, _I, _I, _P, _F, _I, _P, _F, _P],
    'b200ov_input_to_nhwc': [_P, _I, _P, _I, _I, _I, _I, _I, _P, _F, _I, _P, _F, _P],
    'b200ov_input_to_nhwc_split': [_P, _I, _P, _I, _I, _I, _I, _P, _F, _I, _P, _F, _P],
    'b200ov_widen': [_P, _I, _P, _L, _P],
    'b200ov_copy2d': [_P, _P, _L, _I, _I, _I, _P],
    'b200ov_concat_rows': [_I, _P, _P, _P, _L, _P],
    'b200ov_detection_output': [C.POINTER(DetectionDesc), _P, _P, _P, _P, _P],
    'b200ov_detection_output_workspace': [C.POINTER(DetectionDesc), C.POINTER(_Z)],
    'b200ov_detection_output_ws': [C.POINTER(DetectionDesc), _P, _P, _P, _P, _P, _Z, _P],
}
NON_STATUS = {'b200ov_version': ([], _I), 'b200ov_last_error': ([], C.c_char_p)}

_lib = None
launch_count = 0          # kernels launched through this binding (bench.py reports it)

_LAUNCHING = {'b200ov_pack_conv_weights', 'b200ov_conv2d', 'b200ov_conv2d_multi', 'b200ov_matmul', 'b200ov_matmul_ws', 'b200ov_pack_dw_weights', 'b200ov_dwconv2d',
              'b200ov_pool2d', 'b200ov_affine_act', 'b200ov_binary', 'b200ov_softmax', 'b200ov_lrn', 'b200ov_lrn_st', 'b200ov_transpose', 'b200ov_transpose_st', 'b200ov_copy2d_st',
              'b200ov_nchw_to_nhwc_affine', 'b200ov_input_to_nhwc', 'b200ov_input_to_nhwc_split', 'b200ov_widen', 'b200ov_copy2d', 'b200ov_concat_rows', 'b200ov_detection_output', 'b200ov_detection_output_ws'}


def load():
    """dlopen libb200ov.so and declare every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B200ovError('libb200ov.so is not built ({}); run `python -m pyopenvino_b200.build`. '
                          'pyopenvino_b200 has no CPU fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _I
    for name, (args, res) in NON_STATUS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def last_error():
    return load().b200ov_last_error().decode('utf-8', 'replace')


def call(name, *args):
    """Invoke a status-returning entry point; non-zero status becomes a Python exception."""
    global launch_count
    rc = getattr(load(), name)(*args)
    if rc != OK:
        err = B200ovError('{} failed (code {}): {}'.format(name, rc, last_error()))
        err.code = rc
        raise err
    if name in _LAUNCHING:
        launch_count += 1
    return rc
