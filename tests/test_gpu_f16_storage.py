"""GPU parity of the opt-in FP16 STORAGE mode (`load_network(..., storage='f16')`, SURVEY.md 8(f) NEXT-4).

What the mode changes: NHWC feature maps between nodes live in HBM as IEEE half.  What it does not change: every kernel
computes in FP32 (FP32 accumulation; the contraction takes an FP16 input as an exact operand and splits only the
weights), inputs / weights / biases / 2-D tensors / results stay FP32.  Tolerance statement, per op class, against the
oracle evaluated on the SAME (FP16-representable) inputs:

  * MaxPool (selection)                     : bit-exact (max of halfs is a half)
  * Convolution, depthwise, LRN, AvgPool    : |got - ref| <= 2^-11 |ref| (one round-to-nearest-even of the stored result)
                                              + the FP32 class 1e-5 + 1e-4 |ref|   ==>  rtol 6e-4, atol 2e-5
  * end to end                              : rounding accumulates over the stored maps (22 contraction layers deep in
                                              GoogLeNet): classifier probabilities within rtol 2e-2 of the FP32 engine and
                                              the oracle, same top-1 wherever the oracle's top-2 margin exceeds 2 %.
The default storage ('f32') keeps the north-star tolerance; this mode never replaces it.
"""
import os

import numpy as np
import pytest

from conftest import REPO, close

pytestmark = pytest.mark.gpu

RTOL, ATOL = 6e-4, 2e-5


def _to_f16_map(kernels, x_host):
    """Host NCHW array (FP16-representable values) -> NHWC DeviceArray stored as FP16."""
    n, c, h, w = x_host.shape
    x32 = kernels.to_nhwc(kernels.upload(x_host))
    return kernels._into(x32, kernels.new_nhwc(n, c, h, w, st='f16'))


def _rt(x):
    return x.astype(np.float16).astype(np.float32)


@pytest.fixture()
def k16():
    from pyopenvino_b200 import device as dev, kernels
    dev.init()
    prev = kernels.storage
    kernels.storage = 'f16'
    yield kernels
    kernels.storage = prev


@pytest.mark.parametrize('cin,cout,k,s,hw', [(64, 192, 3, 1, 28), (192, 64, 1, 1, 28), (16, 32, 5, 1, 14), (832, 384, 1, 1, 7),
                                              (32, 40, 3, 2, 33), (512, 273, 1, 1, 19), (8, 8, 3, 1, 9)])
def test_conv_f16_in_f16_out_vs_oracle(k16, cin, cout, k, s, hw):
    from oracle import ref_ops
    rng = np.random.default_rng(cin + cout + k)
    n, pad = 3, k // 2
    x = _rt(np.maximum(rng.standard_normal((n, cin, hw, hw)), 0).astype(np.float32))
    w = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    b = (0.05 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)
    want = np.maximum(ref_ops.conv_special(x, w, (s, s), (pad, pad), (pad, pad), 'explicit') + b, 0)
    oh, ow = want.shape[2:]
    xd = _to_f16_map(k16, x)
    assert xd.st == 'f16'
    y = k16.conv2d(xd, k16.upload(w), (s, s), (pad, pad), (oh, ow), bias=k16.upload(b), act=('relu',))
    assert y.st == ('f16' if cout % 8 == 0 else 'f32')
    ok, msg = close(np.asarray(y), want, rtol=RTOL, atol=ATOL)
    assert ok, msg
    # FP32 input -> FP16 output (the stem / first tcgen05 layer of a network) and FP16 input -> FP32 output
    x32 = k16.to_nhwc(k16.upload(x))
    y2 = k16.conv2d(x32, k16.upload(w), (s, s), (pad, pad), (oh, ow), bias=k16.upload(b), act=('relu',))
    ok, msg = close(np.asarray(y2), want, rtol=RTOL, atol=ATOL)
    assert ok, msg
    out32 = k16.new_nhwc(n, cout, oh, ow, st='f32')
    y3 = k16.conv2d(xd, k16.upload(w), (s, s), (pad, pad), (oh, ow), bias=k16.upload(b), act=('relu',), out=out32)
    ok, msg = close(np.asarray(y3), want)            # FP32 out: the north-star tolerance
    assert ok, msg


def test_conv_f16_slices_and_group(k16):
    """FP16 maps as channel slices of a Concat buffer (in and out) and the grouped 1x1 contraction."""
    from oracle import ref_ops
    rng = np.random.default_rng(5)
    n, cin, hw = 2, 192, 28
    x = _rt(np.maximum(rng.standard_normal((n, cin, hw, hw)), 0).astype(np.float32))
    wide = np.zeros((n, cin + 64, hw, hw), np.float32)
    wide[:, 32:32 + cin] = x
    xs = k16.channel_slice(_to_f16_map(k16, wide), 32, cin)
    couts = (64, 96, 16)
    ws = [(rng.standard_normal((co, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32) for co in couts]
    bs = [(0.05 * rng.standard_normal((1, co, 1, 1))).astype(np.float32) for co in couts]
    concat = k16.new_nhwc(n, 256, hw, hw)
    assert concat.st == 'f16'
    members = [(k16.upload(ws[0]), k16.upload(bs[0]), k16.channel_slice(concat, 64, 64)),
               (k16.upload(ws[1]), k16.upload(bs[1]), None), (k16.upload(ws[2]), k16.upload(bs[2]), None)]
    outs = k16.conv1x1_group(xs, members, act=('relu',))
    for w, b, y in zip(ws, bs, outs):
        want = np.maximum(ref_ops.conv_special(x, w, (1, 1), (0, 0), (0, 0), 'explicit') + b, 0)
        ok, msg = close(np.asarray(y), want, rtol=RTOL, atol=ATOL)
        assert ok, msg
    assert outs[0].ld == 256 and outs[0].c_off == 64 and outs[0].st == 'f16'


@pytest.mark.parametrize('hw,k,s,p,rounding,c', [(28, 3, 1, 1, 'ceil', 192), (57, 3, 2, 0, 'ceil', 64), (14, 3, 2, 0, 'ceil', 832), (7, 3, 1, 1, 'ceil', 832),
                                                (28, 2, 2, 0, 'floor', 64), (19, 3, 1, 0, 'floor', 36)])
def test_maxpool_f16_bit_exact(k16, hw, k, s, p, rounding, c):
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi
    rng = np.random.default_rng(hw + c)
    x = _rt((rng.standard_normal((5, c, hw, hw)) * 2 - 1.0).astype(np.float32))
    data = {'strides': '{0},{0}'.format(s), 'kernel': '{0},{0}'.format(k), 'pads_begin': '{0},{0}'.format(p),
            'pads_end': '{0},{0}'.format(p), 'rounding_type': rounding, 'auto_pad': 'explicit'}
    want = ref_ops.maxpool(data, x)
    y = k16.pool2d(_to_f16_map(k16, x), _cabi.POOL_MAX, (k, k), (s, s), (p, p), (p, p), want.shape[2:])
    assert y.st == 'f16'
    assert np.array_equal(np.asarray(y), want)


def test_depthwise_lrn_avgpool_f16_vs_oracle(k16):
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi
    rng = np.random.default_rng(9)
    for hw, s, c in ((38, 1, 256), (75, 2, 128), (10, 1, 1024), (19, 2, 512)):
        x = _rt(rng.standard_normal((3, c, hw, hw)).astype(np.float32))
        w = (rng.standard_normal((c, 1, 1, 3, 3)) * 0.5).astype(np.float32)
        b = (0.1 * rng.standard_normal((1, c, 1, 1))).astype(np.float32)
        want = np.clip(ref_ops.groupconv_numpy(x, w, (s, s), (1, 1), (1, 1), 'explicit') + b, 0.0, 6.0)
        y = k16.dwconv2d(_to_f16_map(k16, x), k16.upload(w), (s, s), (1, 1), want.shape[2:], bias=k16.upload(b), act=('clamp', 0.0, 6.0))
        assert y.st == 'f16'
        ok, msg = close(np.asarray(y), want, rtol=RTOL, atol=ATOL)
        assert ok, ((hw, s, c), msg)
    x = _rt(rng.standard_normal((2, 192, 28, 28)).astype(np.float32))
    data = {'alpha': '9.9999997473787516e-05', 'beta': '0.75', 'bias': '1', 'size': '5'}
    y = k16.lrn(_to_f16_map(k16, x), 5, 9.9999997473787516e-05, 0.75, 1.0)
    assert y.st == 'f16'
    ok, msg = close(np.asarray(y), ref_ops.lrn(data, x), rtol=RTOL, atol=ATOL)
    assert ok, msg
    x = _rt(rng.standard_normal((4, 1024, 7, 7)).astype(np.float32))
    data = {'strides': '1,1', 'kernel': '7,7', 'pads_begin': '0,0', 'pads_end': '0,0', 'rounding_type': 'ceil', 'auto_pad': 'explicit', 'exclude-pad': 'true'}
    want = ref_ops.avgpool(data, x)
    y = k16.pool2d(_to_f16_map(k16, x), _cabi.POOL_AVG_REF, (7, 7), (1, 1), (0, 0), (0, 0), want.shape[2:])
    ok, msg = close(np.asarray(y), want, rtol=RTOL, atol=ATOL)
    assert ok, msg


def _load(model_dir, model, batch, **kw):
    from pyopenvino_b200.inference_engine import IECore
    ie = IECore()
    path = os.path.join(model_dir, model + '.xml')
    net = ie.read_network(path, path[:-4] + '.bin')
    return net, ie.load_network(net, 'B200', batch_size=batch, **kw)


@pytest.mark.parametrize('model', ['googlenet-v1', 'mnist_bn'])
def test_classifiers_f16_storage_end_to_end(model_dir, model):
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    batch = 4
    x = synth_input(model, batch=batch, seed=17)
    net, exe16 = _load(model_dir, model, batch, storage='f16')
    _, exe32 = _load(model_dir, model, batch)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    p16 = exe16.infer({name: x})[out]
    assert np.array_equal(p16, exe16.infer({name: x})[out])               # replay deterministic
    p32 = exe32.infer({name: x})[out]
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    want = oracle.infer_batch(name, x)[out]
    ok, msg = close(p32, want, rtol=1e-4, atol=1e-6)
    assert ok, msg
    for ref in (p32, want):
        ok, msg = close(p16, ref, rtol=2e-2, atol=1e-6)
        assert ok, (model, msg)
    top2 = np.sort(want, axis=1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > 0.02 * top2[:, 1]
    assert np.array_equal(np.argmax(p16, axis=1)[decided], np.argmax(want, axis=1)[decided])
    assert np.allclose(p16.sum(axis=1), 1.0, atol=1e-5)
    # the mode really stores halfs: the working set shrinks
    if model == 'googlenet-v1':
        assert exe16._arena.peak_bytes() < 0.7 * exe32._arena.peak_bytes()
    # uint8 host input + FP16 storage: a second captured graph over the same arena
    x8 = np.clip(np.rint(x * 3), 0, 255).astype(np.uint8)
    a = exe16.infer({name: x8})[out]
    b = exe32.infer({name: x8})[out]
    ok, msg = close(a, b, rtol=2e-2, atol=1e-6)
    assert ok, msg
    assert np.array_equal(exe16.infer({name: x})[out], p16)


def test_ssd_f16_storage_end_to_end(model_dir):
    """SSD-MobileNet with FP16 feature maps: the detection set of the FP32 engine is reproduced -- same classes, boxes
    within 2e-2 (normalised coordinates), scores within 2e-2 relative -- for every record whose score is not within 2 %
    of a neighbour's (near-ties may swap ranks under FP16 rounding; DetectionOutput itself runs in FP32)."""
    from tools.synth_bin import synth_input
    model = 'ssd_mobilenet_v1_coco'
    x = synth_input(model, batch=2, seed=3)
    net, exe16 = _load(model_dir, model, 2, storage='f16')
    _, exe32 = _load(model_dir, model, 2)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    r16 = exe16.infer({name: x})[out][0, 0]
    r32 = exe32.infer({name: x})[out][0, 0]
    for img in range(2):
        a, b = r16[img * 100:(img + 1) * 100], r32[img * 100:(img + 1) * 100]
        na = int(np.where(a[:, 0] == -1)[0][0]) if (a[:, 0] == -1).any() else 100
        nb = int(np.where(b[:, 0] == -1)[0][0]) if (b[:, 0] == -1).any() else 100
        assert nb > 5 and abs(na - nb) <= max(2, nb // 10), (na, nb)
        matched = 0
        for rec in b[:nb]:
            cand = a[:na][a[:na, 1] == rec[1]]
            if len(cand) and np.min(np.abs(cand[:, 3:] - rec[3:]).max(axis=1)) <= 2e-2:
                j = int(np.argmin(np.abs(cand[:, 3:] - rec[3:]).max(axis=1)))
                assert abs(cand[j, 2] - rec[2]) <= 2e-2 * abs(rec[2]) + 1e-4
                matched += 1
        assert matched >= 0.9 * nb, (matched, nb)
    assert exe16._arena.peak_bytes() < 0.75 * exe32._arena.peak_bytes()


def test_f16_storage_range_overflow_falls_back(model_dir):
    """An activation that cannot be STORED as FP16 raises the status word; the engine repeats the inference in FP32 storage
    with the FP32-range kernels and returns finite results."""
    from tools.synth_bin import synth_input
    x = synth_input('mnist_bn', batch=2, seed=5)
    net, exe = _load(model_dir, 'mnist_bn', 2, storage='f16')
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    exe.infer({name: x})
    assert getattr(exe, 'range_fallbacks', 0) == 0
    big = x.copy()
    big[:, :, 10:14, 10:14] = 2.0e5
    got = exe.infer({name: big})[out]
    assert exe.range_fallbacks == 1 and np.all(np.isfinite(got))
    _, exe32 = _load(model_dir, 'mnist_bn', 2, use_graph=False)
    exe32.kernel_type = 'safe'
    assert np.array_equal(got, exe32.infer({name: big})[out])
