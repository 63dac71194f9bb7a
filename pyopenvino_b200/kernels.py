"""Thin Python wrappers: DeviceArray in, DeviceArray out, one libb200ov call each.

This is the only module that talks to the C ABI for compute.  Nothing here does arithmetic on the
host; numpy is used to stage host inputs for upload and nothing else.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from . import device as dev
from .device import DeviceArray

_ACTS = {None: _cabi.ACT_NONE, 'relu': _cabi.ACT_RELU, 'clamp': _cabi.ACT_CLAMP, 'sigmoid': _cabi.ACT_SIGMOID}
default_math = _cabi.MATH_AUTO
# Storage type new NHWC feature maps get: 'f32' (the reference's precision, default) or 'f16' (opt-in storage mode, set by
# the executor for networks loaded with storage='f16').  Arithmetic is FP32 in both; see DeviceArray.st.
storage = 'f32'


def pick_st(c, c_off=0):
    """Storage type for a new feature map of `c` channels: FP16 needs 8-channel (16-byte) pixel granularity."""
    return 'f16' if storage == 'f16' and c % 8 == 0 and c_off % 8 == 0 else 'f32'


def as_f32(x):
    """NHWC feature map in float32 storage (a widening copy when it is stored as FP16): the way into kernels that have no
    FP16-storage variant."""
    if x.layout != 'nhwc' or x.st == 'f32':
        return x
    if x.st == 'hl':
        raise _cabi.B200ovError('an (hi, lo)-pair feature map can only be read by a contraction (planner error)')
    n, c, h, w = x.shape
    out = new_nhwc(n, c, h, w, st='f32')
    _cabi.call('b200ov_copy2d_st', _p(x), x.code, _p(out), out.code, x.pixels, c, x.ld, out.ld, _s())
    return out


def _into(tmp, out):
    """Result `tmp` computed in a temporary -> the preallocated `out` (other storage type / a Concat slot)."""
    if out is None or out is tmp:
        return tmp
    _cabi.call('b200ov_copy2d_st', _p(tmp), tmp.code, _p(out), out.code, tmp.pixels, tmp.shape[1], tmp.ld, out.ld, _s())
    return out


def _act(act):
    """None | ('relu',) | ('clamp', lo, hi) | ('sigmoid',) -> (code, lo, hi)"""
    if act is None:
        return _cabi.ACT_NONE, 0.0, 0.0
    kind = act[0]
    if kind == 'clamp':
        return _cabi.ACT_CLAMP, float(act[1]), float(act[2])
    return _ACTS[kind], 0.0, 0.0


def _p(x):
    """device pointer (or NULL) as c_void_p"""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, DeviceArray):
        return C.c_void_p(x.ptr)
    return C.c_void_p(int(x))


def _s():
    return C.c_void_p(dev.stream())


# ---- range status of the f16x2 contractions -----------------------------------------------------
_status_host = None


def status_reset():
    """Clear the library's sticky status word on the current stream (start of an inference)."""
    _cabi.call('b200ov_status_reset', _s())


def status_fetch(into=None):
    """Queue a D2H copy of the status word into pinned memory (`into`: a pinned int32 tensor, default the shared
    one read by `status_value()`); valid after a stream sync."""
    global _status_host
    if into is None:
        if _status_host is None:
            _status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        into = _status_host
    _cabi.call('b200ov_status_fetch', C.c_void_p(into.data_ptr()), _s())


def status_value():
    return 0 if _status_host is None else int(_status_host[0])


# ---- host <-> device ------------------------------------------------------------------------

def upload(arr, keep_host=False):
    """Host ndarray -> plain DeviceArray (one H2D copy on the current stream)."""
    dev.init()
    a = np.ascontiguousarray(arr, dtype=np.float32)
    t = dev.alloc_f32(a.size)
    if a.size:
        t[:a.size].copy_(torch.from_numpy(a.reshape(-1)), non_blocking=False)
    out = DeviceArray(t, a.shape, 'plain')
    if keep_host:
        out.cache['host'] = a
    return out


def as_device(x):
    return x if isinstance(x, DeviceArray) else upload(x)


def host_value(x):
    """Small constant operand as a host ndarray (no device sync when the Const plugin cached it)."""
    if isinstance(x, DeviceArray):
        if 'host' not in x.cache:
            x.cache['host'] = x.numpy()
        return x.cache['host']
    return np.asarray(x)


def to_nhwc(x, scale=None, shift=None, out=None):
    """plain NCHW -> NHWC (optionally y = x*scale[c] + shift[c] on the fly: fused Parameter pre-processing)."""
    x = as_device(x)
    assert x.layout == 'plain' and x.ndim == 4
    n, c, h, w = x.shape
    if out is None:
        # a 3-channel network input gets a pixel pitch of 4 floats (zero-filled pad lane) so the stem
        # convolution can gather 16-byte runs and run on the tensor cores (conv_f16x2.cu, "pair" mode)
        ld = 4 if c == 3 else c
        out = DeviceArray(dev.alloc_f32(n * h * w * ld), x.shape, 'nhwc', ld=ld)
    sv, ss, hs = _affine_operand(scale, c)
    bv, bs, hb = _affine_operand(shift, c)
    _cabi.call('b200ov_nchw_to_nhwc_affine', _p(x), _p(out), n, c, h * w, out.ld, hs, sv, ss, hb, bv, bs, _s())
    return out


def input_to_device(raw, scale=None, shift=None, nhwc=True, split=False):
    """RawInput (native element type, IR layout) -> float32 DeviceArray: NHWC (+ folded scale / shift) for 4-D inputs
    when `nhwc`, else a plain widened copy.  One kernel; the widening is exact."""
    if raw.ndim == 4 and nhwc:
        n, c, h, w = raw.shape
        ld = 4 if c == 3 else c
        sv, ss, hs = _affine_operand(scale, c)
        bv, bs, hb = _affine_operand(shift, c)
        if split and c <= 4:
            # the stem Convolution is the only consumer: write the FP16 (hi, lo) pairs of the contraction here, once per pixel
            out = DeviceArray(dev.alloc_f32(n * h * w * 4), raw.shape, 'nhwc', ld=4, st='hl')
            _cabi.call('b200ov_input_to_nhwc_split', C.c_void_p(raw.ptr), raw.code, _p(out), n, c, h * w, hs, sv, ss, hb, bv, bs, _s())
            return out
        out = DeviceArray(dev.alloc_f32(n * h * w * ld), raw.shape, 'nhwc', ld=ld)
        _cabi.call('b200ov_input_to_nhwc', C.c_void_p(raw.ptr), raw.code, _p(out), n, c, h * w, out.ld, hs, sv, ss, hb, bv, bs, _s())
        return out
    assert scale is None and shift is None
    out = DeviceArray(dev.alloc_f32(raw.size), raw.shape, 'plain')
    _cabi.call('b200ov_widen', C.c_void_p(raw.ptr), raw.code, _p(out), raw.size, _s())
    return out


def upload_raw(arr):
    """Host ndarray of a native input type -> RawInput (one H2D copy of arr.nbytes on the current stream)."""
    dev.init()
    a = np.ascontiguousarray(dev.native_input(arr))
    t = torch.empty(max(a.size, 1), dtype=dev.RawInput.TORCH[a.dtype], device='cuda')
    if a.size:
        t[:a.size].copy_(torch.from_numpy(a.reshape(-1)), non_blocking=False)
    return dev.RawInput(t, a.shape, a.dtype)


def to_plain(x):
    """NHWC -> plain NCHW (the layout a host consumer or a flattening Reshape needs)."""
    assert x.layout == 'nhwc'
    n, c, h, w = x.shape
    out = DeviceArray(dev.alloc_f32(n * c * h * w), x.shape, 'plain')
    if h * w == 1 and x.is_dense():
        _cabi.call('b200ov_copy2d_st', _p(x), x.code, _p(out), _cabi.DT_F32, n, c, x.ld, c, _s())
    else:
        _cabi.call('b200ov_transpose_st', _p(x), x.code, _p(out), _cabi.DT_F32, n, h * w, c, x.ld, h * w, _s())
    return out


def as_nhwc(x):
    x = as_device(x)
    if x.layout == 'nhwc':
        return x
    assert x.ndim == 4, 'expected a 4-D feature map, got shape {}'.format(x.shape)
    return to_nhwc(x)


def as_plain(x):
    x = as_device(x)
    return to_plain(x) if x.layout == 'nhwc' else x


def new_nhwc(n, c, h, w, st=None):
    st = pick_st(c) if st is None else st
    elems = n * h * w * c
    return DeviceArray(dev.alloc_f32((elems + 1) // 2 if st == 'f16' else elems), (n, c, h, w), 'nhwc', ld=c, st=st)


def _check_out(out, shape):
    assert out.layout == 'nhwc' and tuple(out.shape) == tuple(shape), \
        'preallocated output {} does not match result shape {}'.format(out, shape)
    return out


# ---- operands of the elementwise tail ----------------------------------------------------------

def _affine_operand(b, c):
    """-> (vector pointer or NULL, scalar value, has_flag) for a broadcast operand of `c` channels."""
    if b is None:
        return C.c_void_p(0), 0.0, 0
    size = b.size
    if size == 1:
        return C.c_void_p(0), float(host_value(b).reshape(-1)[0]), 1
    if size == c:
        d = as_device(b)
        assert d.layout == 'plain'
        return C.c_void_p(d.ptr), 0.0, 1
    raise _cabi.B200ovError('unsupported broadcast: operand with {} elements against {} channels'.format(size, c))


def _vec_ptr(b, c):
    if b is None:
        return None
    d = as_device(b)
    assert d.layout == 'plain' and d.size == c, 'per-channel operand must have {} elements, got {}'.format(c, d.size)
    return d


# ---- Convolution / MatMul ----------------------------------------------------------------------

class PackedWeights:
    __slots__ = ('t', 'rows', 'ldw', 'cout', 'cin', 'kh', 'kw')

    @property
    def ptr(self):
        return self.t.data_ptr()


def pack_conv(w):
    """OIHW (or [N][K] for MatMul) -> packed [kh*kw*cin][ldw]; cached on the weight array."""
    w = as_device(w)
    if 'conv' in w.cache:
        return w.cache['conv']
    assert w.layout == 'plain'
    if w.ndim == 4:
        cout, cin, kh, kw = w.shape
    else:
        cout, cin = w.shape
        kh = kw = 1
    rows, ldw, total = C.c_int(0), C.c_int(0), C.c_int64(0)
    _cabi.call('b200ov_conv_weight_dims', cout, cin, kh, kw, C.byref(rows), C.byref(ldw), C.byref(total))
    pk = PackedWeights()
    # packed weights are long-lived: never from the per-inference arena
    pk.t = torch.empty(total.value, dtype=torch.float32, device='cuda')
    pk.rows, pk.ldw, pk.cout, pk.cin, pk.kh, pk.kw = rows.value, ldw.value, cout, cin, kh, kw
    _cabi.call('b200ov_pack_conv_weights', _p(w), C.c_void_p(pk.ptr), cout, cin, kh, kw, _s())
    w.cache['conv'] = pk
    return pk


def f16x2_ok(x, cin, act_code, mode):
    """The library's eligibility rule for the f16x2 contraction (conv_f16x2.cu: f16x2_eligible) -- the only kernel that reads
    or writes FP16-stored feature maps."""
    if mode not in (_cabi.MATH_AUTO, _cabi.MATH_F16X2) or act_code == _cabi.ACT_SIGMOID or x.ptr % 16 != 0:
        return False
    if x.st == 'f16':
        return x.ld % 8 == 0 and cin % 8 == 0
    if x.st == 'hl':
        return True                      # produced for exactly this consumer (inference_engine.build_plan checks the geometry)
    return x.ld % 4 == 0 and (cin % 8 == 0 or (cin <= 4 and x.ld == 4))


def conv2d(x, w, strides, pads_begin, out_hw, bias=None, act=None, out=None, math=None, pre_pool=False, hl_out=False):
    """`hl_out`: every reader of the result is a contraction -- where the f16x2 kernel takes this layer it writes the (hi, lo)
    pair form those readers would otherwise compute per filter tap (DeviceArray.st == 'hl'; same bits downstream).
    `pre_pool`: x is the INPUT of a MaxPool 3x3 / stride 1 / pads 1 node whose only consumer is this 1x1 convolution; the
    pooling runs inside the contraction's A producers (b200ov_conv_desc.pre_pool) where the f16x2 kernel takes the layer,
    as a separate b200ov_pool2d otherwise (FP32-range re-run, FP16 storage, unaligned slices)."""
    x = as_nhwc(x)
    pk = w if isinstance(w, PackedWeights) else pack_conv(w)
    n, c, h, wd = x.shape
    assert c == pk.cin, 'input has {} channels, filter expects {}'.format(c, pk.cin)
    oh, ow = out_hw
    shape = (n, pk.cout, oh, ow)
    code, lo, hi = _act(act)
    mode = default_math if math is None else math
    if pre_pool:
        fusable = (x.st == 'f32' and f16x2_ok(x, c, code, mode) and c % 8 == 0 and (pk.kh, pk.kw) == (1, 1) and
                   tuple(strides) == (1, 1) and tuple(pads_begin) == (0, 0) and (oh, ow) == (h, wd) and wd <= 63 and
                   (out is None or out.st == 'f32') and pick_st(pk.cout) == 'f32')
        if not fusable:
            x = pool2d(x, _cabi.POOL_MAX, (3, 3), (1, 1), (1, 1), (1, 1), (h, wd))
            pre_pool = False
    final = _check_out(out, shape) if out is not None else None
    if x.st == 'hl' and mode not in (_cabi.MATH_AUTO, _cabi.MATH_F16X2):
        raise _cabi.B200ovError('a pre-split network input can only feed the f16x2 contraction')
    if not f16x2_ok(x, c, code, mode):
        x = as_f32(x)                    # the FP32-range kernels read and write float32 feature maps only
    half_ok = f16x2_ok(x, c, code, mode)
    if final is not None and (final.st == 'f32' or half_ok):
        out = final
    else:
        st = pick_st(pk.cout) if half_ok else 'f32'
        # C_in = 1 stems run the direct FP32 kernel, which can write the pair form too (conv_ffma.cu: conv2d_c1_direct_ok)
        c1_hl = (c == 1 and x.st == 'f32' and x.ld == 1 and (pk.kh, pk.kw) in ((3, 3), (5, 5)) and pk.cout <= 64 and
                 (pk.cout // 4) & (pk.cout // 4 - 1) == 0 and mode == _cabi.MATH_AUTO and code != _cabi.ACT_SIGMOID and
                 storage == 'f32' and os.environ.get('B200OV_NO_C1_DIRECT') is None)
        if hl_out and (half_ok or c1_hl) and st == 'f32' and final is None and pk.cout % 8 == 0:
            st = 'hl'
        out = new_nhwc(*shape, st=st)
    d = _cabi.ConvDesc(n=n, h=h, w=wd, cin=c, cout=pk.cout, kh=pk.kh, kw=pk.kw, sh=strides[0], sw=strides[1],
                       pt=pads_begin[0], pl=pads_begin[1], oh=oh, ow=ow, x_ld=x.ld, y_ld=out.ld, ldw=pk.ldw,
                       act=code, act_lo=lo, act_hi=hi, math=mode, x_dtype=x.code, y_dtype=out.code,
                       pre_pool=_cabi.PREPOOL_MAX3X3S1 if pre_pool else _cabi.PREPOOL_NONE)
    b = _vec_ptr(bias, pk.cout)
    _cabi.call('b200ov_conv2d', C.byref(d), _p(x), C.c_void_p(pk.ptr), _p(b), _p(out), _s())
    return _into(out, final)


def conv1x1_group(x, members, act=None):
    """Sibling 1x1 convolutions of one feature map as a single contraction (b200ov_conv2d_multi).
    members = [(w OIHW DeviceArray, bias DeviceArray | None, out DeviceArray | None), ...] (2 or 3).
    Returns the list of outputs.  The fused weight matrix (members concatenated along C_out, each starting at a
    multiple of 32 columns) is built and packed once and cached on the first member's weight."""
    x = as_nhwc(x)
    n, c, h, wd = x.shape
    hl_flags = [bool(m[3]) if len(m) > 3 else False for m in members]
    members = [tuple(m[:3]) for m in members]
    ws = [as_device(m[0]) for m in members]
    key = ('group',) + tuple(id(w.t) for w in ws)
    cache = ws[0].cache
    if cache.get('group_key') != key:
        col0s, total = [], 0
        for w in ws:
            assert w.ndim == 4 and w.shape[1] == c and w.shape[2] == 1 and w.shape[3] == 1
            col0s.append(total)
            total += (w.shape[0] + 31) // 32 * 32
        fw = torch.zeros(total * c, dtype=torch.float32, device='cuda')
        fb = torch.zeros(total, dtype=torch.float32, device='cuda')
        for (w, (_, b, _o), col0) in zip(ws, members, col0s):
            co = w.shape[0]
            fw.view(total, c)[col0:col0 + co].copy_(w.t[:co * c].view(co, c))
            if b is not None:
                bd = as_device(b)
                fb[col0:col0 + co].copy_(bd.t[:co])
        pk = pack_conv(DeviceArray(fw, (total, c, 1, 1), 'plain'))
        cache['group_key'] = key
        cache['group'] = (pk, DeviceArray(fb, (total,), 'plain'), col0s, total)
    pk, fb, col0s, total = cache['group']
    outs, finals = [], []
    # one storage type for all members: FP16 only when every member (and every preallocated slot) allows it
    group_st = 'f16' if all((m[2].st == 'f16') if m[2] is not None else pick_st(w.shape[0]) == 'f16' for w, m in zip(ws, members)) else 'f32'
    if x.st == 'f16' and not (x.ld % 8 == 0 and x.ptr % 16 == 0):
        x = as_f32(x)
    segs = (_cabi.ConvSeg * len(members))()
    for i, (w, (_, _b, o)) in enumerate(zip(ws, members)):
        shape = (n, w.shape[0], h, wd)
        final = _check_out(o, shape) if o is not None else None
        if final is None and hl_flags[i] and group_st == 'f32' and w.shape[0] % 8 == 0:
            o = new_nhwc(*shape, st='hl')               # read by contractions only: written in their operand form
        else:
            o = final if final is not None and final.st == group_st else new_nhwc(*shape, st=group_st)
        finals.append(final)
        outs.append(o)
        segs[i].y = o.ptr
        segs[i].col0, segs[i].cout, segs[i].y_ld = col0s[i], w.shape[0], o.ld
        segs[i].y_dtype = _cabi.DT_HL if o.st == 'hl' else _cabi.DT_F32
    code, lo, hi = _act(act)
    st = group_st
    assert all(o.st in (st, 'hl') for o in outs), 'grouped convolution outputs must share one storage type'
    d = _cabi.ConvDesc(n=n, h=h, w=wd, cin=c, cout=total, kh=1, kw=1, sh=1, sw=1, pt=0, pl=0, oh=h, ow=wd, x_ld=x.ld, y_ld=total,
                       ldw=pk.ldw, act=code, act_lo=lo, act_hi=hi, math=_cabi.MATH_AUTO, x_dtype=x.code,
                       y_dtype=_cabi.DT_F16 if group_st == 'f16' else _cabi.DT_F32)
    _cabi.call('b200ov_conv2d_multi', C.byref(d), _p(x), C.c_void_p(pk.ptr), _p(fb), len(members), segs, _s())
    return [_into(o, f) for o, f in zip(outs, finals)]


def matmul(a, b, transpose_a=False, transpose_b=True, bias=None, act=None, math=None):
    a = as_plain(a)
    assert a.ndim == 2
    if transpose_a:
        k, m = a.shape
        at = DeviceArray(dev.alloc_f32(m * k), (m, k), 'plain')
        _cabi.call('b200ov_transpose', _p(a), _p(at), 1, k, m, m, k, _s())
        a = at
    m, k = a.shape
    b = as_device(b)
    key = 'mm_tb' if transpose_b else 'mm_plain'
    if key not in b.cache:
        if transpose_b:
            nk = b                          # already [N][K]
        else:
            kk, nn = b.shape                # [K][N] -> [N][K]
            nk = DeviceArray(torch.empty(kk * nn, dtype=torch.float32, device='cuda'), (nn, kk), 'plain')
            _cabi.call('b200ov_transpose', _p(b), _p(nk), 1, kk, nn, nn, kk, _s())
        b.cache[key] = pack_conv(nk)
    pk = b.cache[key]
    assert pk.cin == k, 'MatMul inner dims differ: {} vs {}'.format(k, pk.cin)
    n = pk.cout
    out = DeviceArray(dev.alloc_f32(m * n), (m, n), 'plain')
    code, lo, hi = _act(act)
    bv = _vec_ptr(bias, n)
    mode = default_math if math is None else math
    need = C.c_size_t(0)
    if mode in (_cabi.MATH_AUTO, _cabi.MATH_F16X2):
        _cabi.call('b200ov_matmul_workspace', m, n, k, C.byref(need))
    if need.value:
        # few output tiles: split-K over a workspace of partial sums (from the arena, like every per-inference buffer)
        ws = dev.alloc_f32(need.value // 4)
        _cabi.call('b200ov_matmul_ws', m, n, k, _p(a), k, C.c_void_p(pk.ptr), pk.ldw, _p(bv), code, lo, hi, mode, _p(out), n,
                   C.c_void_p(ws.data_ptr()), need, _s())
        _cabi.launch_count += 1          # two kernels: the contraction and the reduction
    else:
        _cabi.call('b200ov_matmul', m, n, k, _p(a), k, C.c_void_p(pk.ptr), pk.ldw, _p(bv), code, lo, hi, mode, _p(out), n, _s())
    return out


# ---- depthwise ---------------------------------------------------------------------------------

def pack_dw(w):
    w = as_device(w)
    if 'dw' in w.cache:
        return w.cache['dw']
    g, co, ci, kh, kw = w.shape
    if co != 1 or ci != 1:
        raise _cabi.B200ovError('GroupConvolution: only depthwise (C_out/G = C_in/G = 1) is supported, like the reference')
    t = torch.empty(g * kh * kw, dtype=torch.float32, device='cuda')
    _cabi.call('b200ov_pack_dw_weights', _p(w), C.c_void_p(t.data_ptr()), g, kh, kw, _s())
    w.cache['dw'] = (t, g, kh, kw)
    return w.cache['dw']


def dwconv2d(x, w, strides, pads_begin, out_hw, bias=None, act=None, out=None, exact=False, hl_out=False):
    """`hl_out`: every reader of the result is a contraction (depthwise -> pointwise): written in their (hi, lo) operand form
    where the tile kernel takes the layer (DeviceArray.st == 'hl'), as FP32 otherwise."""
    x = as_nhwc(x)
    t, g, kh, kw = pack_dw(w)
    n, c, h, wd = x.shape
    assert c == g, 'depthwise: {} input channels vs {} groups'.format(c, g)
    oh, ow = out_hw
    shape = (n, c, oh, ow)
    code, lo, hi = _act(act)
    final = _check_out(out, shape) if out is not None else None
    # FP16-stored maps: the TMA tile kernel (3x3, stride 1 / 2, type-preserving); anything else goes through float32
    half = x.st == 'f16' and not exact and kh == 3 and kw == 3 and strides[0] == strides[1] and strides[0] in (1, 2) and \
        code <= _cabi.ACT_CLAMP and c % 4 == 0 and x.ld % 8 == 0 and x.ptr % 16 == 0
    if not half:
        x = as_f32(x)
    if final is not None and final.st == x.st and (x.st == 'f32' or (final.ld % 4 == 0 and final.ptr % 8 == 0)):
        out = final
    else:
        out = new_nhwc(*shape, st=x.st)
    d = _cabi.DwConvDesc(n=n, h=h, w=wd, c=c, kh=kh, kw=kw, sh=strides[0], sw=strides[1], pt=pads_begin[0],
                         pl=pads_begin[1], oh=oh, ow=ow, x_ld=x.ld, y_ld=out.ld, act=code, act_lo=lo, act_hi=hi,
                         math=_cabi.DW_EXACT if exact else _cabi.DW_AUTO, dtype=x.code)
    b = _vec_ptr(bias, c)
    if hl_out and final is None and x.st == 'f32' and not exact and c % 8 == 0 and storage == 'f32':
        out_hl = new_nhwc(*shape, st='hl')
        d.y_dtype = _cabi.DT_HL
        try:
            _cabi.call('b200ov_dwconv2d', C.byref(d), _p(x), C.c_void_p(t.data_ptr()), _p(b), _p(out_hl), _s())
            return out_hl
        except _cabi.B200ovError as e:
            if e.code != _cabi.ERR_UNSUPPORTED:
                raise
            d.y_dtype = 0                      # no pair-writing kernel for this shape: plain FP32 output
    _cabi.call('b200ov_dwconv2d', C.byref(d), _p(x), C.c_void_p(t.data_ptr()), _p(b), _p(out), _s())
    return _into(out, final)


# ---- pooling -----------------------------------------------------------------------------------

def pool2d(x, mode, kernel, strides, pads_begin, pads_end, out_hw, scale=None, shift=None, out=None):
    x = as_nhwc(x)
    n, c, h, wd = x.shape
    oh, ow = out_hw
    shape = (n, c, oh, ow)
    final = _check_out(out, shape) if out is not None else None
    if x.st == 'f16' and not (c % 4 == 0 and x.ld % 4 == 0 and x.ptr % 8 == 0):
        x = as_f32(x)
    if final is not None and final.st == x.st and (x.st == 'f32' or (final.ld % 4 == 0 and final.ptr % 8 == 0)):
        out = final
    else:
        out = new_nhwc(*shape, st=x.st)                    # pooling keeps the storage type of its input
    d = _cabi.PoolDesc(n=n, h=h, w=wd, c=c, kh=kernel[0], kw=kernel[1], sh=strides[0], sw=strides[1],
                       pt=pads_begin[0], pl=pads_begin[1], pb=pads_end[0], pr=pads_end[1], oh=oh, ow=ow,
                       x_ld=x.ld, y_ld=out.ld, mode=mode, dtype=x.code)
    _cabi.call('b200ov_pool2d', C.byref(d), _p(x), _p(_vec_ptr(scale, c)), _p(_vec_ptr(shift, c)), _p(out), _s())
    return _into(out, final)


# ---- elementwise tail ----------------------------------------------------------------------------

def _rows_channels(x):
    """(rows, channels, x_ld) view of an activation for channel-broadcast elementwise kernels."""
    if x.layout == 'nhwc':
        return x.pixels, x.shape[1], x.ld
    c = x.shape[-1] if x.ndim >= 1 else 1
    return x.size // max(c, 1), c, c


def _like(x, out=None):
    if out is not None:
        return _check_out(out, x.shape)
    if x.layout == 'nhwc':
        n, c, h, w = x.shape
        return new_nhwc(n, c, h, w, st=x.st)
    return DeviceArray(dev.alloc_f32(x.size), x.shape, 'plain')


def _channel_operand_ok(x, b):
    """True when `b` broadcasts against `x` as a scalar or a per-channel vector in x's physical layout."""
    if b.size == 1:
        return True
    bs = tuple(b.shape)
    if x.layout == 'nhwc':
        c = x.shape[1]
        return b.size == c and len(bs) == 4 and bs[1] == c
    # plain: channel = last dim
    return b.size == x.shape[-1] and bs[-1] == x.shape[-1]


def affine_act(x, scale=None, shift=None, act=None, out=None):
    """act(x*scale + shift) with scalar / per-channel operands (either may be None)."""
    x = as_f32(as_device(x))             # standalone elementwise nodes: float32 storage only (fused ones never get here)
    if x.layout == 'plain' and x.ndim == 4 and ((scale is not None and scale.size > 1) or (shift is not None and shift.size > 1)):
        x = to_nhwc(x)
    rows, c, x_ld = _rows_channels(x)
    final = out
    if out is not None and out.st != 'f32':
        out = None
    out = _like(x, out)
    y_ld = out.ld if out.layout == 'nhwc' else c
    sv, ss, hs = _affine_operand(scale, c)
    bv, bs, hb = _affine_operand(shift, c)
    code, lo, hi = _act(act)
    _cabi.call('b200ov_affine_act', _p(x), _p(out), rows, c, x_ld, y_ld, hs, sv, ss, hb, bv, bs, code, lo, hi, _s())
    return _into(out, final) if final is not None and final is not out else out


def binary(op, a, b):
    """Same-shape Add (op 0) / Multiply (op 1)."""
    a, b = as_f32(as_device(a)), as_f32(as_device(b))
    assert tuple(a.shape) == tuple(b.shape)
    if a.layout != b.layout or not a.is_dense() or not b.is_dense():
        a, b = as_plain(a), as_plain(b)
    out = _like(a)
    _cabi.call('b200ov_binary', op, _p(a), _p(b), _p(out), a.size, _s())
    return out


def softmax_rows(x):
    x = as_plain(x)
    rows = x.shape[0] if x.ndim >= 2 else 1
    cols = x.size // rows
    out = DeviceArray(dev.alloc_f32(x.size), x.shape, 'plain')
    _cabi.call('b200ov_softmax', _p(x), _p(out), rows, cols, _s())
    return out


def lrn(x, size, alpha, beta, bias, out=None):
    x = as_nhwc(x)
    c = x.shape[1]
    if x.st == 'f16' and not (c % 4 == 0 and x.ld % 4 == 0 and x.ptr % 8 == 0 and size // 2 <= 4):
        x = as_f32(x)
    final = out
    if out is not None and (out.st != x.st or (x.st == 'f16' and not (out.ld % 4 == 0 and out.ptr % 8 == 0))):
        out = None
    out = _like(x, out)
    _cabi.call('b200ov_lrn_st', _p(x), _p(out), x.code, x.pixels, c, x.ld, out.ld, size, alpha, beta, bias, _s())
    return _into(out, final) if final is not None and final is not out else out


def copy_channels(src, dst):
    """Copy an NHWC array into an NHWC channel slice of equal logical shape."""
    assert src.layout == 'nhwc' and dst.layout == 'nhwc' and tuple(src.shape) == tuple(dst.shape)
    _cabi.call('b200ov_copy2d_st', _p(src), src.code, _p(dst), dst.code, src.pixels, src.shape[1], src.ld, dst.ld, _s())
    return dst


def channel_slice(buf, c_off, c):
    """View of channels [c_off, c_off + c) of an NHWC buffer."""
    n, ct, h, w = buf.shape
    assert buf.layout == 'nhwc' and c_off + c <= ct
    return DeviceArray(buf.t, (n, c, h, w), 'nhwc', ld=buf.ld, c_off=buf.c_off + c_off, st=buf.st)


def detection_output(loc, conf, proposals, num_classes, keep_top_k, center_size, variance_in_target, clip_before, clip_after,
                     conf_thr, nms_thr):
    """SSD post-process for a batch: loc [N, P*4], conf [N, P*classes], proposals [1, 2, P*4] -> [1, 1, N*keep, 7]."""
    loc, conf, proposals = as_plain(loc), as_plain(conf), as_plain(proposals)
    n = loc.shape[0]
    priors = proposals.shape[2] // 4
    assert loc.size == n * priors * 4 and conf.size == n * priors * num_classes
    out = DeviceArray(dev.alloc_f32(n * keep_top_k * 7), (1, 1, n * keep_top_k, 7), 'plain')
    d = _cabi.DetectionDesc(n=n, num_priors=priors, num_classes=num_classes, keep_top_k=keep_top_k,
                            code_center_size=int(center_size), variance_encoded_in_target=int(variance_in_target),
                            clip_before_nms=int(clip_before), clip_after_nms=int(clip_after),
                            confidence_threshold=conf_thr, nms_threshold=nms_thr)
    # scratch for the batch-wide top-1 pass: (score, class) per prior
    nbytes = C.c_size_t(0)
    _cabi.call('b200ov_detection_output_workspace', C.byref(d), C.byref(nbytes))
    ws = dev.alloc_f32((nbytes.value + 3) // 4)
    _cabi.call('b200ov_detection_output_ws', C.byref(d), _p(loc), _p(conf), _p(proposals), _p(out), C.c_void_p(ws.data_ptr()),
               C.c_size_t(nbytes.value), _s())
    _cabi.launch_count += 1              # two kernels: the batch-wide top-1 pass and the per-image CTAs
    return out
