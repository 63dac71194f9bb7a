"""CPU-side checks of the C-ABI boundary: libb200ov.so builds for sm_100a, loads, and exports every
symbol include/b200ov.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import REPO


@pytest.fixture(scope='module')
def lib_path():
    from pyopenvino_b200 import build
    return build.build()


def _declared():
    text = open(os.path.join(REPO, 'include', 'b200ov.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200ov_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_hot_path():
    names = _declared()
    for must in ('b200ov_conv2d', 'b200ov_dwconv2d', 'b200ov_matmul', 'b200ov_pool2d', 'b200ov_affine_act',
                 'b200ov_softmax', 'b200ov_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), 'libb200ov.so does not export ' + name


def test_binding_covers_every_declared_symbol(lib_path):
    from pyopenvino_b200 import _cabi
    bound = set(_cabi.SIGNATURES) | set(_cabi.NON_STATUS)
    assert bound == set(_declared())
    lib = _cabi.load()
    assert lib.b200ov_version() == 100
    # descriptor layouts must match the C structs (all 4-byte fields)
    assert ctypes.sizeof(_cabi.ConvDesc) == 4 * 23
    assert ctypes.sizeof(_cabi.DwConvDesc) == 4 * 20
    assert ctypes.sizeof(_cabi.PoolDesc) == 4 * 18


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_invalid_descriptor_is_rejected_without_a_gpu(lib_path):
    """Validation happens before any CUDA call, so error reporting can be checked on the CPU box."""
    from pyopenvino_b200 import _cabi
    lib = _cabi.load()
    d = _cabi.ConvDesc()           # all zeros: invalid
    rc = lib.b200ov_conv2d(ctypes.byref(d), None, None, None, None, None)
    assert rc == _cabi.ERR_INVALID
    assert b'conv2d' in lib.b200ov_last_error()
    with pytest.raises(_cabi.B200ovError):
        _cabi.call('b200ov_softmax', None, None, 1, 10, None)


def test_round2_entry_points_validate_without_a_gpu(lib_path):
    """b200ov_concat_rows / b200ov_detection_output_ws (round 2): argument checks and the workspace size are host-side."""
    from pyopenvino_b200 import _cabi
    lib = _cabi.load()
    assert lib.b200ov_concat_rows(0, None, None, None, 1, None) == _cabi.ERR_INVALID
    assert b'concat_rows' in lib.b200ov_last_error()
    srcs = (ctypes.c_void_p * 9)(*([None] * 9))
    cols = (ctypes.c_int * 9)(*([4] * 9))
    assert lib.b200ov_concat_rows(9, srcs, cols, None, 1, None) == _cabi.ERR_INVALID        # more than B200OV_CONCAT_MAX_PARTS
    d = _cabi.DetectionDesc(n=64, num_priors=1917, num_classes=91, keep_top_k=100)
    nbytes = ctypes.c_size_t(0)
    assert lib.b200ov_detection_output_workspace(ctypes.byref(d), ctypes.byref(nbytes)) == 0
    assert nbytes.value == 64 * 1917 * 8                      # (score, class) per prior
    assert lib.b200ov_detection_output_ws(ctypes.byref(d), None, None, None, None, None, 0, None) == _cabi.ERR_INVALID
