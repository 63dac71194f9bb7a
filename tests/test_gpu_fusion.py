"""GPU parity of the producer-side fusions of the contraction kernel (round 2): the 3x3 / stride-1 MaxPool folded into
the A producers of the 1x1 convolution that consumes it (b200ov_conv_desc.pre_pool; MaxPool.py:41-72 + Convolution.py:57-87)."""
import os

import numpy as np
import pytest

from conftest import close

pytestmark = pytest.mark.gpu

POOL_DATA = {'strides': '1,1', 'kernel': '3,3', 'pads_begin': '1,1', 'pads_end': '1,1', 'rounding_type': 'ceil', 'auto_pad': 'explicit'}


@pytest.mark.parametrize('shape', [
    # n, cin, h, w, cout, act, sliced input / output
    (3, 192, 28, 28, 32, ('relu',), False),       # inception_3a/pool -> pool_proj
    (2, 480, 14, 14, 64, ('relu',), True),        # inception_4a, input = a channel slice of a wider buffer, output into a Concat slot
    (5, 832, 7, 7, 128, ('relu',), False),        # inception_5a
    (2, 528, 14, 14, 128, None, True),         # inception_4e
    (1, 24, 5, 9, 40, None, False),            # odd map, C_out not a multiple of 32, K shorter than a slot
    (2, 16, 1, 1, 8, ('relu',), False),           # 1x1 image: the window is all padding but the centre
    (1, 8, 3, 63, 16, ('clamp', -0.5, 0.75), False),   # widest map the pixel-tile box covers (128 + 2w + 2 <= 256 rows)
    (1, 8, 3, 200, 16, ('clamp', -0.5, 0.75), False),  # wider: kernels.conv2d issues the two kernels (1 + 1 launches)
    (7, 40, 1, 33, 72, None, False),           # single image row
    (2, 64, 33, 1, 96, ('relu',), False),         # single image column
])
def test_pool_fused_conv_bit_identical_to_separate_kernels_and_vs_oracle(shape):
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, cin, h, w, cout, act, sliced = shape
    rng = np.random.default_rng(abs(hash(shape[:5])) % (1 << 31))
    # mixed signs: at the image border an all-negative window must produce 0 (the zero padding takes part in the max)
    x = (rng.standard_normal((n, cin, h, w)) - 0.8).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)
    if sliced:
        wide = np.full((n, cin + 16, h, w), np.float32(1e9))
        wide[:, 8:8 + cin] = x
        xd = kernels.channel_slice(kernels.to_nhwc(kernels.upload(wide)), 8, cin)
        out_f = kernels.channel_slice(kernels.new_nhwc(n, cout + 64, h, w), 32, cout)
        out_s = kernels.channel_slice(kernels.new_nhwc(n, cout + 64, h, w), 32, cout)
    else:
        xd = kernels.to_nhwc(kernels.upload(x))
        out_f = out_s = None
    wd, bd = kernels.upload(wt), kernels.upload(b)
    kernels.pack_conv(wd)
    launches0 = _cabi.launch_count
    fused = kernels.conv2d(xd, wd, (1, 1), (0, 0), (h, w), bias=bd, act=act, out=out_f, pre_pool=True)
    assert _cabi.launch_count - launches0 == (1 if w <= 63 else 2), 'the fused pair must be ONE library call'
    pooled = kernels.pool2d(xd, _cabi.POOL_MAX, (3, 3), (1, 1), (1, 1), (1, 1), (h, w))
    separate = kernels.conv2d(pooled, wd, (1, 1), (0, 0), (h, w), bias=bd, act=act, out=out_s)
    got, sep = np.asarray(fused), np.asarray(separate)
    assert np.array_equal(got, sep), 'fused pool->conv differs from pool2d + conv2d (max |d| = {})'.format(np.abs(got - sep).max())
    p = ref_ops.maxpool(POOL_DATA, x)
    assert np.array_equal(np.asarray(pooled), p)
    want = ref_ops.conv_special(p, wt, (1, 1), (0, 0), (0, 0), 'explicit') + b
    if act == ('relu',):
        want = np.maximum(want, 0)
    elif act is not None:
        want = np.clip(want, act[1], act[2])
    ok, msg = close(got, want)
    assert ok, msg


def test_pool_fusion_falls_back_outside_the_f16x2_path():
    """pre_pool with the FP32-range arithmetic (the range fallback's re-run) or FP16 storage: kernels.conv2d issues the
    separate MaxPool; the C entry point itself refuses instead of silently dropping the pooling."""
    import ctypes as C
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((2, 32, 9, 9)) - 0.5).astype(np.float32)
    wt = (rng.standard_normal((24, 32, 1, 1)) * 0.2).astype(np.float32)
    xd, wd = kernels.to_nhwc(kernels.upload(x)), kernels.upload(wt)
    want = ref_ops.conv_special(ref_ops.maxpool(POOL_DATA, x), wt, (1, 1), (0, 0), (0, 0), 'explicit')
    for math in (_cabi.MATH_SAFE, _cabi.MATH_FP32, _cabi.MATH_TF32X3):
        y = kernels.conv2d(xd, wd, (1, 1), (0, 0), (9, 9), math=math, pre_pool=True)
        ok, msg = close(np.asarray(y), want)
        assert ok, (math, msg)
    pk = kernels.pack_conv(wd)
    out = kernels.new_nhwc(2, 24, 9, 9)
    d = _cabi.ConvDesc(n=2, h=9, w=9, cin=32, cout=24, kh=1, kw=1, sh=1, sw=1, pt=0, pl=0, oh=9, ow=9, x_ld=xd.ld, y_ld=out.ld,
                       ldw=pk.ldw, act=0, act_lo=0.0, act_hi=0.0, math=_cabi.MATH_FP32, x_dtype=0, y_dtype=0,
                       pre_pool=_cabi.PREPOOL_MAX3X3S1)
    rc = _cabi.load().b200ov_conv2d(C.byref(d), C.c_void_p(xd.ptr), C.c_void_p(pk.ptr), None, C.c_void_p(out.ptr), None)
    assert rc == _cabi.ERR_UNSUPPORTED


def test_googlenet_pool_fusion_end_to_end(model_dir):
    """The nine pool -> pool_proj pairs of GoogLeNet run fused (nine launches fewer per inference) and the network
    output is bit-identical to the plan without the fusion (B200OV_NO_POOL_FUSE=1)."""
    from pyopenvino_b200.inference_engine import IECore
    from tools.synth_bin import synth_input
    ie = IECore()
    xml = os.path.join(model_dir, 'googlenet-v1.xml')
    x = synth_input('googlenet-v1', batch=8, seed=3)
    outs, launches = [], []
    for off in ('0', '1'):
        os.environ['B200OV_NO_POOL_FUSE'] = off
        try:
            net = ie.read_network(xml, xml[:-4] + '.bin')
            exe = ie.load_network(net, 'B200', batch_size=8)
            name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
            outs.append(exe.infer({name: x})[out_name])
            outs.append(exe.infer({name: x})[out_name])          # graph replay
            launches.append(exe.kernels_per_inference())
            if off == '0':
                assert sum(1 for s in exe._plan.values() if s.get('pool_into') is not None) == 9
        finally:
            os.environ.pop('B200OV_NO_POOL_FUSE', None)
    assert launches[1] - launches[0] == 9, launches
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


@pytest.mark.parametrize('shape', [
    # n, cin, hw, cmid, cout, k, stride, pad
    (2, 192, 28, 96, 128, 3, 1, 1),        # inception_3a 3x3_reduce -> 3x3
    (3, 480, 14, 16, 48, 5, 1, 2),         # inception_4a 5x5_reduce -> 5x5
    (2, 64, 19, 128, 256, 3, 2, 1),        # SSD extra layer: 1x1 -> 3x3 / stride 2
    (1, 40, 9, 24, 20, 3, 1, 1),           # cmid = 24: K tail inside a slot, odd map
])
def test_hl_edge_between_contractions_is_bit_identical_and_vs_oracle(shape):
    """A 1x1 convolution whose only reader is another convolution leaves its output as FP16 (hi, lo) pairs
    (B200OV_DT_HL); the reader must produce exactly the bits it produces from the FP32 tensor."""
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, cin, hw, cmid, cout, k, s, pad = shape
    rng = np.random.default_rng(cin * 131 + cout)
    x = np.maximum(rng.standard_normal((n, cin, hw, hw)), 0).astype(np.float32)
    w1 = (rng.standard_normal((cmid, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal((1, cmid, 1, 1))).astype(np.float32)
    w2 = (rng.standard_normal((cout, cmid, k, k)) * np.sqrt(2.0 / (cmid * k * k))).astype(np.float32)
    b2 = (0.1 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)
    xd = kernels.to_nhwc(kernels.upload(x))
    w1d, b1d, w2d, b2d = (kernels.upload(a) for a in (w1, b1, w2, b2))
    oh = (hw + 2 * pad - k) // s + 1
    outs = []
    for hl in (True, False):
        mid = kernels.conv2d(xd, w1d, (1, 1), (0, 0), (hw, hw), bias=b1d, act=('relu',), hl_out=hl)
        assert mid.st == ('hl' if hl else 'f32')
        y = kernels.conv2d(mid, w2d, (s, s), (pad, pad), (oh, oh), bias=b2d, act=('relu',))
        assert y.st == 'f32'
        outs.append((np.asarray(mid), np.asarray(y)))
    assert np.array_equal(outs[0][1], outs[1][1]), 'reader of the (hi, lo) tensor differs from the reader of the FP32 tensor'
    # the debug decode of the pair form carries 22 significant bits of the FP32 values
    assert np.allclose(outs[0][0], outs[1][0], rtol=3e-7, atol=1e-30)
    mid_ref = np.maximum(ref_ops.conv_special(x, w1, (1, 1), (0, 0), (0, 0), 'explicit') + b1, 0)
    want = np.maximum(ref_ops.conv_special(mid_ref, w2, (s, s), (pad, pad), (pad, pad), 'explicit') + b2, 0)
    ok, msg = close(outs[0][1], want)
    assert ok, msg


def test_hl_group_segments(model_dir):
    """b200ov_conv2d_multi with mixed FP32 / (hi, lo) output segments (the 1x1 | 3x3_reduce | 5x5_reduce contraction of an
    inception module): the FP32 segment and the readers of the pair segments are bit-identical to the all-FP32 run."""
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(77)
    n, cin, hw = 2, 256, 14
    couts = (128, 96, 32)
    x = np.maximum(rng.standard_normal((n, cin, hw, hw)), 0).astype(np.float32)
    ws = [kernels.upload((rng.standard_normal((co, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32)) for co in couts]
    bs = [kernels.upload((0.05 * rng.standard_normal((1, co, 1, 1))).astype(np.float32)) for co in couts]
    w3 = kernels.upload((rng.standard_normal((64, 96, 3, 3)) * 0.05).astype(np.float32))
    w5 = kernels.upload((rng.standard_normal((48, 32, 5, 5)) * 0.05).astype(np.float32))
    xd = kernels.to_nhwc(kernels.upload(x))
    res = []
    for hl in (True, False):
        outs = kernels.conv1x1_group(xd, [(ws[0], bs[0], None, False), (ws[1], bs[1], None, hl), (ws[2], bs[2], None, hl)], act=('relu',))
        assert [o.st for o in outs] == (['f32', 'hl', 'hl'] if hl else ['f32'] * 3)
        y3 = kernels.conv2d(outs[1], w3, (1, 1), (1, 1), (hw, hw))
        y5 = kernels.conv2d(outs[2], w5, (1, 1), (2, 2), (hw, hw))
        res.append([np.asarray(outs[0]), np.asarray(y3), np.asarray(y5)])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('model,batch', [('googlenet-v1', 8), ('ssd_mobilenet_v1_coco', 2), ('mnist_bn', 96)])
def test_models_with_hl_edges_bit_identical(model_dir, model, batch):
    """Whole networks: the plan with contraction -> contraction tensors in (hi, lo) form against B200OV_NO_HL=1."""
    from pyopenvino_b200.inference_engine import IECore
    from tools.synth_bin import synth_input
    ie = IECore()
    xml = os.path.join(model_dir, model + '.xml')
    x = synth_input(model, batch=batch, seed=5)
    outs, edges = [], []
    for off in ('0', '1'):
        os.environ['B200OV_NO_HL'] = off
        try:
            net = ie.read_network(xml, xml[:-4] + '.bin')
            exe = ie.load_network(net, 'B200', batch_size=batch)
            name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
            outs.append(exe.infer({name: x})[out_name])
            outs.append(exe.infer({name: x})[out_name])
            edges.append(sum(1 for s in exe._plan.values() if s['ops'].get('hl_out')))
        finally:
            os.environ.pop('B200OV_NO_HL', None)
    assert edges[0] > 0 and edges[1] == 0, edges
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


@pytest.mark.parametrize('shape', [(4, 32, 150, 1, 64), (3, 64, 75, 2, 128), (2, 512, 19, 1, 512), (2, 40, 11, 1, 24)])
def test_hl_edge_depthwise_to_pointwise(shape):
    """Depthwise 3x3 (+bias +Clamp) writing the pointwise convolution's (hi, lo) operand form: the pointwise result must be
    bit-identical to the one computed from the FP32 depthwise output."""
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, c, hw, s, cout = shape
    rng = np.random.default_rng(c * 7 + hw)
    x = np.maximum(rng.standard_normal((n, c, hw, hw)), 0).astype(np.float32)
    wdw = kernels.upload((rng.standard_normal((c, 1, 1, 3, 3)) * 0.3).astype(np.float32))
    bdw = kernels.upload((0.1 * rng.standard_normal((1, c, 1, 1))).astype(np.float32))
    wpw = kernels.upload((rng.standard_normal((cout, c, 1, 1)) * np.sqrt(2.0 / c)).astype(np.float32))
    xd = kernels.to_nhwc(kernels.upload(x))
    oh = (hw + 2 - 3) // s + 1
    res = []
    for hl in (True, False):
        mid = kernels.dwconv2d(xd, wdw, (s, s), (1, 1), (oh, oh), bias=bdw, act=('clamp', 0.0, 6.0), hl_out=hl)
        if not hl:
            assert mid.st == 'f32'
        y = kernels.conv2d(mid, wpw, (1, 1), (0, 0), (oh, oh), act=('clamp', 0.0, 6.0))
        res.append((mid.st, np.asarray(y)))
    assert np.array_equal(res[0][1], res[1][1]), res[0][0]
    if n * c * oh * oh >= (1 << 16):
        assert res[0][0] == 'hl'                # the tile kernel took it; tiny shapes fall back to FP32 (still identical)
