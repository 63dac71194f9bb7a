"""GPU parity: every op plugin of pyopenvino_b200, called through the reference's compute() contract
with HOST arrays (so the call crosses the C ABI with H2D / D2H inside), against vectors the live
reference produced (tests/golden/ops.npz) and against the oracle on fresh seeded inputs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, close

pytestmark = pytest.mark.gpu

OPS = np.load(os.path.join(GOLDEN, 'ops.npz'))
META = json.loads(str(OPS['meta']))

# selection / exact-arithmetic ops must be bit-identical to the reference; the rest use the
# north-star FP32 tolerance |a-b| <= 1e-5 + 1e-4*|b|
BIT_EXACT = {'GroupConvolution', 'MaxPool', 'Add', 'Multiply', 'ReLU', 'Clamp', 'Concat', 'Transpose', 'Reshape',
             'Unsqueeze', 'ShapeOf', 'StridedSlice', 'PriorBoxClustered', 'DetectionOutput'}


@pytest.fixture(scope='module')
def plugins():
    from pyopenvino_b200.inference_engine import IECore
    return IECore().plugins.plugins


def _node(i, m):
    prec = {np.dtype('float32'): 'FP32', np.dtype('int64'): 'I64'}
    ins = {p: OPS['c{}_in{}'.format(i, p)] for p in m['ports']}
    out_port = 1 if m['type'] == 'ShapeOf' else len(ins)
    node = {'name': m['tag'], 'type': m['type'], 'data': dict(m['data']),
            'input': {p: {'precision': prec[a.dtype], 'dims': tuple(a.shape)} for p, a in ins.items()},
            'output': {out_port: {'precision': 'I64' if m['type'] == 'ShapeOf' else 'FP32', 'dims': ()}}}
    return node, ins, out_port


@pytest.mark.parametrize('i', range(len(META)), ids=[m['tag'] for m in META])
def test_plugin_vs_reference_vectors(i, plugins):
    m = META[i]
    node, ins, op = _node(i, m)
    want = OPS['c{}_out_numpy'.format(i)]
    if m['type'] == 'GroupConvolution':
        # default kernel = packed-FMA chain (tolerance class); 'exact' = pairwise kernel, bit-identical (checked below)
        fast = np.asarray(plugins[m['type']].compute(node, dict(ins), kernel_type='numpy')[op])
        ok, msg = close(fast, want)
        assert ok, (m['tag'], msg)
    res = plugins[m['type']].compute(node, dict(ins), kernel_type='exact' if m['type'] == 'GroupConvolution' else 'numpy')
    got = np.asarray(res[op])
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype
    if m['type'] in BIT_EXACT:
        assert np.array_equal(got, want), '{}: max |d| = {}'.format(m['tag'], np.abs(got - want).max())
    else:
        ok, msg = close(got, want)
        assert ok, (m['tag'], msg)


@pytest.mark.parametrize('kt', ['fp32', 'tf32x3', 'tf32', 'f16x2'])
def test_conv_math_modes(kt, plugins):
    """'fp32' = CUDA-core FFMA, 'tf32x3' = tcgen05 hi/lo split (both must meet the FP32 tolerance),
    'tf32' = single-pass tcgen05: only the relaxed 1e-3 class, checked relative to the tensor scale;
    'f16x2' = persistent tcgen05 kernel, FP16 hi/lo split with the A operand in TMEM (FP32 tolerance)."""
    from pyopenvino_b200 import _cabi
    ran = 0
    for i, m in enumerate(META):
        if m['type'] != 'Convolution':
            continue
        node, ins, op = _node(i, m)
        want = OPS['c{}_out_numpy'.format(i)]
        try:
            got = np.asarray(plugins['Convolution'].compute(node, dict(ins), kernel_type=kt)[op])
        except _cabi.B200ovError as e:
            assert kt != 'fp32' and ('tcgen05 path needs' in str(e) or 'f16x2 path needs' in str(e)), str(e)   # C_in = 1 / 3 stems
            continue
        ran += 1
        if kt == 'tf32':
            assert np.abs(got - want).max() <= 2e-3 * max(1.0, np.abs(want).max()), m['tag']
        else:
            ok, msg = close(got, want)
            assert ok, (m['tag'], kt, msg)
    assert ran >= 5


def test_conv_pickle_known_answer(plugins):
    """resources/node_args_6.pickle (SSD conv0, real weights): f32 crop vs the reference's output."""
    g = np.load(os.path.join(GOLDEN, 'conv_kat.npz'))
    data = json.loads(str(g['node']))['data']
    x, w = g['x_f32_crop'], g['w_f32']
    node = {'name': 'conv0', 'type': 'Convolution', 'data': data,
            'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
            'output': {2: {'precision': 'FP32', 'dims': (1, 32, 32, 32)}}}
    got = plugins['Convolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2]
    ok, msg = close(got, g['y_f32_crop_numpy'])
    assert ok, msg
    # full 300x300 image, f16-quantised real data up-cast to f32, against the oracle
    from oracle import ref_ops
    x, w = g['x_f16'].astype(np.float32), g['w_f16'].astype(np.float32)
    node['input'] = {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}}
    got = plugins['Convolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2]
    want = ref_ops.convolution(data, x, w, 'special')
    ok, msg = close(got, want)
    assert ok, msg


def _conv_node(x, w, s, pb, pe, auto_pad='explicit'):
    data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
            'pads_end': '{}, {}'.format(*pe), 'auto_pad': auto_pad}
    return {'name': 'conv', 'type': 'Convolution', 'data': data,
            'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
            'output': {2: {'precision': 'FP32', 'dims': ()}}}


@pytest.mark.parametrize('shape', [
    # (N, Cin, H, Cout, k, s, pads_begin, pads_end): GoogLeNet / SSD / MNIST layer shapes, batched
    (2, 64, 56, 192, 3, 1, (1, 1), (1, 1)), (3, 192, 28, 64, 1, 1, (0, 0), (0, 0)), (2, 16, 28, 32, 5, 1, (2, 2), (2, 2)),
    (2, 3, 224, 64, 7, 2, (3, 3), (3, 3)), (2, 832, 7, 384, 1, 1, (0, 0), (0, 0)), (2, 3, 300, 32, 3, 2, (0, 0), (1, 1)),
    (1, 512, 19, 273, 1, 1, (0, 0), (0, 0)), (2, 256, 10, 512, 3, 2, (1, 1), (1, 1)), (4, 1, 28, 64, 3, 1, (1, 1), (1, 1)),
    (2, 528, 14, 256, 1, 1, (0, 0), (0, 0)), (1, 160, 14, 320, 3, 1, (1, 1), (1, 1)), (5, 24, 9, 12, 1, 1, (0, 0), (0, 0)),
    # C_in <= 4 stems: two-half pair gather (stride 1 / odd width) and the super-pixel mapping (even stride and width) with
    # even and odd left padding, even and odd kernel widths
    (2, 3, 33, 16, 3, 1, (1, 1), (1, 1)), (2, 3, 31, 8, 3, 2, (1, 1), (1, 1)), (2, 3, 32, 16, 5, 2, (2, 2), (2, 2)),
    (1, 4, 30, 8, 4, 2, (1, 1), (2, 2)), (2, 2, 20, 40, 2, 2, (0, 0), (0, 0)),
    # 96-column tiles (C_out in (64, 96], (128, 192], (256, 288]) next to 128-column ones, odd slot counts
    (2, 64, 14, 288, 3, 1, (1, 1), (1, 1)), (2, 32, 14, 96, 1, 1, (0, 0), (0, 0)), (1, 96, 9, 160, 3, 1, (1, 1), (1, 1))])
def test_conv_layer_shapes_vs_oracle(shape, plugins):
    from oracle import ref_ops
    n, cin, hw, cout, k, s, pb, pe = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    x = rng.standard_normal((n, cin, hw, hw)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    node = _conv_node(x, w, s, pb, pe)
    got = plugins['Convolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2]
    want = ref_ops.conv_special(x, w, (s, s), pb, pe, 'explicit')      # 'special' handles N > 1
    ok, msg = close(got, want)
    assert ok, (shape, msg)


@pytest.mark.parametrize('shape', [(2, 32, 150, 1, (1, 1), (1, 1)), (2, 64, 150, 2, (0, 0), (1, 1)), (3, 128, 75, 2, (1, 1), (1, 1)),
                                   (2, 512, 19, 1, (1, 1), (1, 1)), (2, 1024, 10, 1, (1, 1), (1, 1)), (1, 6, 7, 1, (1, 1), (1, 1)),
                                   (2, 8, 9, 2, (0, 0), (1, 1)), (1, 16, 21, 2, (1, 1), (1, 1)), (2, 4, 5, 1, (1, 1), (1, 1))])
def test_depthwise_bit_exact_vs_oracle(shape, plugins):
    from oracle import ref_ops
    n, c, hw, s, pb, pe = shape
    rng = np.random.default_rng(c * 7 + hw)
    x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
    w = (rng.standard_normal((c, 1, 1, 3, 3)) * np.sqrt(2.0 / 9)).astype(np.float32)
    data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
            'pads_end': '{}, {}'.format(*pe), 'auto_pad': 'explicit'}
    node = {'name': 'dw', 'type': 'GroupConvolution', 'data': data,
            'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
            'output': {2: {'precision': 'FP32', 'dims': ()}}}
    want = ref_ops.groupconv_numpy(x, w, (s, s), pb, pe, 'explicit')
    got = plugins['GroupConvolution'].compute(node, {0: x, 1: w}, kernel_type='exact')[2]
    assert np.array_equal(got, want), np.abs(got - want).max()        # pairwise kernel: bit-identical to np.sum
    fast = plugins['GroupConvolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2]
    ok, msg = close(fast, want)                                         # packed-FMA 3x3 kernel: FP32 tolerance class
    assert ok, (shape, msg)
    assert np.abs(fast - want).max() <= 4e-6, np.abs(fast - want).max()


@pytest.mark.parametrize('shape', [(2, 64, 112, 3, 2, (0, 0), 'ceil'), (2, 192, 28, 3, 1, (1, 1), 'ceil'), (3, 832, 7, 3, 1, (1, 1), 'ceil'),
                                   (4, 32, 26, 2, 2, (0, 0), 'floor'), (2, 10, 11, 2, 2, (0, 0), 'floor'),
                                   (2, 8, 19, 3, 2, (1, 1), 'ceil'), (1, 12, 23, 3, 2, (0, 0), 'floor'), (2, 4, 9, 2, 1, (0, 0), 'floor'),
                                   (1, 16, 17, 3, 1, (0, 0), 'floor')])
def test_maxpool_bit_exact_vs_oracle(shape, plugins):
    from oracle import ref_ops
    n, c, hw, k, s, p, rounding = shape
    rng = np.random.default_rng(c + hw)
    x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)      # negative values: zero padding matters
    data = {'strides': '{}, {}'.format(s, s), 'kernel': '{}, {}'.format(k, k), 'pads_begin': '{}, {}'.format(*p),
            'pads_end': '{}, {}'.format(*p), 'rounding_type': rounding, 'auto_pad': 'explicit'}
    node = {'name': 'pool', 'type': 'MaxPool', 'data': data, 'input': {0: {'precision': 'FP32', 'dims': x.shape}},
            'output': {1: {'precision': 'FP32', 'dims': ()}}}
    got = plugins['MaxPool'].compute(node, {0: x}, kernel_type='numpy')[1]
    assert np.array_equal(got, ref_ops.maxpool(data, x))


def test_matmul_batched_rows_vs_oracle(plugins):
    from oracle import ref_ops
    rng = np.random.default_rng(3)
    for (m, k, n) in [(1, 1024, 1000), (64, 6272, 512), (7, 64, 10), (33, 128, 10)]:
        a = rng.standard_normal((m, k)).astype(np.float32)
        b = (rng.standard_normal((n, k)) * np.sqrt(2.0 / k)).astype(np.float32)
        data = {'transpose_a': 'false', 'transpose_b': 'true'}
        node = {'name': 'mm', 'type': 'MatMul', 'data': data,
                'input': {0: {'precision': 'FP32', 'dims': a.shape}, 1: {'precision': 'FP32', 'dims': b.shape}},
                'output': {2: {'precision': 'FP32', 'dims': ()}}}
        got = plugins['MatMul'].compute(node, {0: a, 1: b}, kernel_type='numpy')[2]
        ok, msg = close(got, ref_ops.matmul(data, a, b))
        assert ok, ((m, k, n), msg)


def test_softmax_rows_are_independent_images(plugins):
    from oracle import ref_ops
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((6, 1000)) * 2).astype(np.float32)
    node = {'name': 'sm', 'type': 'SoftMax', 'data': {'axis': '1'}, 'input': {0: {'precision': 'FP32', 'dims': x.shape}},
            'output': {1: {'precision': 'FP32', 'dims': ()}}}
    got = plugins['SoftMax'].compute(node, {0: x}, kernel_type='numpy')[1]
    want = np.concatenate([ref_ops.softmax(x[i:i + 1]) for i in range(6)])
    ok, msg = close(got, want, rtol=1e-4, atol=1e-7)
    assert ok, msg
    assert np.array_equal(np.argmax(got, axis=1), np.argmax(want, axis=1))


def test_contract_violation_raises(plugins):
    x = np.zeros((1, 3, 8, 8), dtype=np.float32)
    node = {'name': 'relu', 'type': 'ReLU', 'data': {}, 'input': {0: {'precision': 'FP32', 'dims': (1, 3, 9, 9)}},
            'output': {1: {'precision': 'FP32', 'dims': ()}}}
    with pytest.raises(AssertionError):
        plugins['ReLU'].compute(node, {0: x})
    with pytest.raises(AssertionError):
        node['input'][0]['dims'] = (1, 3, 8, 8)
        plugins['ReLU'].compute(node, {0: x.astype(np.float64)})


def test_f16x2_range_overflow_falls_back_to_full_range(plugins):
    """An activation beyond the FP16 range (65504) makes the f16x2 contraction raise the status word; the
    plugin (host-in / host-out) then repeats the node with the FP32-range kernels, so the result still
    matches the reference.  With kernel_type='f16x2' pinned, the overflow is visible as a non-finite output."""
    from oracle import ref_ops
    rng = np.random.default_rng(7)
    x = rng.standard_normal((1, 16, 9, 9)).astype(np.float32)
    x[0, 3, 4, 4] = 3.0e5
    w = (rng.standard_normal((24, 16, 3, 3)) * 0.1).astype(np.float32)
    node = _conv_node(x, w, 1, (1, 1), (1, 1))
    want = ref_ops.conv_special(x, w, (1, 1), (1, 1), (1, 1), 'explicit')
    got = plugins['Convolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2]
    assert np.all(np.isfinite(got))
    ok, msg = close(got, want, rtol=1e-4, atol=1e-5 * 3.0e5)       # absolute slack scaled to the 3e5 activation
    assert ok, msg
    pinned = plugins['Convolution'].compute(node, {0: x, 1: w}, kernel_type='f16x2')[2]
    assert not np.all(np.isfinite(pinned))


def test_sibling_1x1_group_matches_separate_convs(plugins):
    """b200ov_conv2d_multi: three 1x1 convolutions of one feature map as a single contraction, one of them
    writing into a channel slice of a wider (Concat) buffer; every output must match the oracle."""
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    rng = np.random.default_rng(11)
    n, cin, hw = 3, 192, 28
    couts = (64, 96, 16)
    x = np.maximum(rng.standard_normal((n, cin, hw, hw)), 0).astype(np.float32)
    ws = [(rng.standard_normal((co, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32) for co in couts]
    bs = [(0.05 * rng.standard_normal((1, co, 1, 1))).astype(np.float32) for co in couts]
    xd = kernels.to_nhwc(kernels.upload(x))
    concat = kernels.new_nhwc(n, 256, hw, hw)
    members = [(kernels.upload(ws[0]), kernels.upload(bs[0]), kernels.channel_slice(concat, 64, 64)),
               (kernels.upload(ws[1]), kernels.upload(bs[1]), None), (kernels.upload(ws[2]), kernels.upload(bs[2]), None)]
    outs = kernels.conv1x1_group(xd, members, act=('relu',))
    for w, b, y in zip(ws, bs, outs):
        want = np.maximum(ref_ops.conv_special(x, w, (1, 1), (0, 0), (0, 0), 'explicit') + b, 0)
        ok, msg = close(y.numpy(), want)
        assert ok, msg
    assert outs[0].ld == 256 and outs[0].c_off == 64


@pytest.mark.parametrize('tool,cases', [('fuzz_conv.py', 200), ('fuzz_ops.py', 240)])
def test_randomised_parity_sweeps(tool, cases):
    """tools/fuzz_conv.py / fuzz_ops.py: random shapes, strides, paddings and fused epilogues against the oracle."""
    import subprocess
    import sys
    from conftest import REPO
    r = subprocess.run([sys.executable, os.path.join(REPO, 'tools', tool), '--cases', str(cases), '--seed', '11'],
                       capture_output=True, text=True, cwd=REPO, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
