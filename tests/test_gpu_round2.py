"""GPU parity, round 2: native-width host inputs, the async slot API, fused-plan tensors against the oracle,
the f16x2 contraction on wide-dynamic-range data, and the INTEGRATION.md option-B binding through raw ctypes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO, close

pytestmark = pytest.mark.gpu


def _load(model_dir, model, batch=None, fuse=True, use_graph=True, **kw):
    from pyopenvino_b200.inference_engine import IECore
    ie = IECore()
    path = os.path.join(REPO, 'models', 'mnist.xml') if model == 'mnist' else os.path.join(model_dir, model + '.xml')
    net = ie.read_network(path, path[:-4] + '.bin')
    exe = ie.load_network(net, 'B200', batch_size=batch, fuse=fuse, use_graph=use_graph, **kw)
    return net, exe


def _records(block):
    stop = np.where(block[:, 0] == -1)[0]
    return block[:int(stop[0])] if len(stop) else block


# ---- native-width host inputs (Parameter.py:13: any array-like, cast with .astype(precision)) ----------------------

@pytest.mark.parametrize('use_graph', [True, False])
def test_mnist_uint8_image_like_draw_and_infer(model_dir, use_graph):
    """draw-and-infer.py:56-60 hands a uint8 28x28 canvas to infer(); README.md:69-72 known answer."""
    g = np.load(os.path.join(GOLDEN, 'mnist_e2e.npz'))
    img_u8 = g['input'].reshape(28, 28).astype(np.uint8)
    assert np.array_equal(img_u8.astype(np.float32), g['input'].reshape(28, 28))
    net, exe = _load(model_dir, 'mnist', use_graph=use_graph)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    for _ in range(2):
        prob = exe.infer({name: img_u8})[out]
        assert list(np.argsort(prob[0])[::-1]) == [2, 0, 1, 7, 8, 6, 3, 4, 5, 9]
        ok, msg = close(prob, g['final_numpy'], rtol=1e-4, atol=1e-7)
        assert ok, msg
    # the same pixels as float32, float16 and a python list: bit-identical results (the widening is exact)
    ref = exe.infer({name: img_u8.astype(np.float32)})[out]
    assert np.array_equal(ref, prob)
    assert np.array_equal(exe.infer({name: img_u8.astype(np.float16)})[out], ref)
    assert np.array_equal(exe.infer({name: img_u8.astype(np.int64)})[out], ref)          # other dtypes: host cast like the reference
    assert np.array_equal(exe.infer({name: img_u8.tolist()})[out], ref)


def test_ssd_uint8_input_vs_oracle(model_dir):
    """SSD-MobileNet on uint8 camera-style frames: 1 byte per pixel crosses PCIe, `Preprocessor/mul` + `/sub` and the
    cast run in the layout kernel.  Records identical in order / class to the oracle fed the same uint8 frames, and
    bit-identical to this engine fed the frames as float32."""
    from oracle import ref_engine
    model = 'ssd_mobilenet_v1_coco'
    rng = np.random.default_rng(11)
    x8 = rng.integers(0, 256, (2, 3, 300, 300), dtype=np.uint8)
    net, exe = _load(model_dir, model, batch=2)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    got8 = exe.infer({name: x8})[out]
    got32 = exe.infer({name: x8.astype(np.float32)})[out]
    assert np.array_equal(got8, got32)
    assert np.array_equal(exe.infer({name: x8})[out], got8)           # switching graphs back and forth is stable
    assert len(exe._captured) == 2
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    for img in range(2):
        want = _records(oracle.infer({name: x8[img:img + 1]})[out][0, 0])
        got = _records(got8[0, 0, img * 100:(img + 1) * 100])
        assert len(want) > 3
        assert got.shape == want.shape
        assert np.array_equal(got[:, 0:2], want[:, 0:2])
        ok, msg = close(got[:, 2:], want[:, 2:], rtol=1e-4, atol=1e-5)
        assert ok, msg


def test_googlenet_float16_and_uint8_inputs_bit_exact(model_dir):
    from tools.synth_bin import synth_input
    x16 = synth_input('googlenet-v1', batch=3, seed=4).astype(np.float16)
    net, exe = _load(model_dir, 'googlenet-v1', batch=3)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    a = exe.infer({name: x16})[out]
    b = exe.infer({name: x16.astype(np.float32)})[out]
    assert np.array_equal(a, b)
    x8 = np.random.default_rng(5).integers(0, 3, x16.shape, dtype=np.uint8)
    assert np.array_equal(exe.infer({name: x8})[out], exe.infer({name: x8.astype(np.float32)})[out])
    # eager, unfused path takes the same inputs
    net2, exe2 = _load(model_dir, 'googlenet-v1', batch=3, fuse=False, use_graph=False)
    c = exe2.infer({name: x16})[out]
    ok, msg = close(a, c, rtol=1e-5, atol=1e-7)
    assert ok, msg


# ---- async slot API -------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('warm', [0, 1, 2, 3])
def test_async_slots_zero_copy_does_not_depend_on_call_parity(model_dir, warm):
    """`start_async(..., slot=s)` with the slot's own `request_buffer(s, ...)` never copies on the host, whatever the
    number of requests issued before; round-robin callers can ask `next_slot()`."""
    from tools.synth_bin import synth_input
    net, exe = _load(model_dir, 'mnist_bn', batch=4)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    xs = [synth_input('mnist_bn', batch=4, seed=40 + i) for i in range(6)]
    want = [exe.infer({name: x})[out] for x in xs]
    for _ in range(warm):
        exe.wait(exe.start_async({name: xs[0]}))
    bufs = [exe.request_buffer(s, name) for s in range(exe.NUM_REQUESTS)]
    got, pending = [], None
    for i, x in enumerate(xs):
        s = exe.next_slot() if i % 2 else i % exe.NUM_REQUESTS
        if pending is not None and pending == s:
            got.append(exe.wait(pending)[out])
            pending = None
        bufs[s][...] = x
        slot = exe.start_async({name: bufs[s]}, slot=s)
        assert slot == s
        if pending is not None:
            got.append(exe.wait(pending)[out])
        pending = slot
    got.append(exe.wait(pending)[out])
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # uint8 request buffers
    b8 = exe.request_buffer(0, name, np.uint8)
    b8[...] = 7
    r8 = exe.wait(exe.start_async({name: b8}, slot=0))[out]
    assert np.array_equal(r8, exe.infer({name: np.full(b8.shape, 7, np.float32)})[out])


# ---- the fused plan against the oracle on tensors SoftMax does not compress -------------------------------------------

def _device_outputs(net):
    got = {}
    for nid in net.G.nodes:
        n = net.G.nodes[nid]
        if 'output' in n:
            p = next(iter(n['output']))
            if 'data' in n['output'][p]:
                got[n['name']] = n['output'][p]['data']
    return got


def test_fused_plan_logits_vs_oracle(model_dir):
    """Fused plan (epilogue fusion, grouped 1x1, in-place Concat), GoogLeNet: the PRE-SoftMax logits and every inception
    output against the oracle's tensors for the same nodes.  Tolerance: |d| <= 1e-5*max(1, max|ref|) + 1e-4*|ref|
    (end to end over 22 contraction layers; the reference's own special-vs-numpy kernels differ by 9e-6 end to end)."""
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    model = 'googlenet-v1'
    x = synth_input(model, batch=1, seed=2)
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    oracle.infer({oracle.net.inputs[0]['name']: x})
    want = oracle.node_outputs()
    net, exe = _load(model_dir, model, batch=1, fuse=True, use_graph=False)
    exe.infer({net.inputs[0]['name']: x})
    got = _device_outputs(net)
    softmax_in = [n for n in oracle.net.nodes.values() if n['type'] == 'SoftMax'][0]
    fl, fp, _ = oracle.net.pred[[k for k, v in oracle.net.nodes.items() if v is softmax_in][0]][0]
    names = [oracle.net.nodes[fl]['name']] + [n['name'] for n in oracle.net.nodes.values() if n['type'] == 'Concat'] + ['pool5/7x7_s1']
    checked = 0
    for nm in names:
        assert nm in got, nm
        w = want[nm]
        ok, msg = close(np.asarray(got[nm]), w, rtol=1e-4, atol=1e-5 * max(1.0, float(np.abs(w).max())))
        assert ok, (nm, msg)
        checked += 1
    assert checked == 11
    assert np.argmax(np.asarray(got[names[0]])) == np.argmax(want[names[0]])


def test_fused_plan_ssd_heads_vs_oracle(model_dir):
    """SSD-MobileNet fused plan: the three DetectionOutput inputs (box deltas, class confidences after Sigmoid, priors)
    and the backbone feature maps feeding the heads, against the oracle."""
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    model = 'ssd_mobilenet_v1_coco'
    x = synth_input(model, batch=1, seed=6)
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    oracle.infer({oracle.net.inputs[0]['name']: x})
    want = oracle.node_outputs()
    net, exe = _load(model_dir, model, batch=1, fuse=True, use_graph=False)
    exe.infer({net.inputs[0]['name']: x})
    got = _device_outputs(net)
    det = [k for k, v in oracle.net.nodes.items() if v['type'] == 'DetectionOutput'][0]
    names = [oracle.net.nodes[fl]['name'] for fl, fp, tp in oracle.net.pred[det]]
    names += [n['name'] for n in oracle.net.nodes.values() if n['type'] == 'Clamp'][-6:]
    for nm in names:
        assert nm in got, nm
        w = want[nm]
        ok, msg = close(np.asarray(got[nm]), w, rtol=1e-4, atol=1e-5 * max(1.0, float(np.abs(w).max())))
        assert ok, (nm, msg)


def test_ssd_batch64_records_vs_oracle(model_dir):
    """configs[3] at the bench size: record blocks of a batch-64 run against the ORACLE run on the same images."""
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    model = 'ssd_mobilenet_v1_coco'
    x = synth_input(model, batch=64, seed=33)
    net, exe = _load(model_dir, model, batch=64)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    full = exe.infer({name: x})[out]
    keep = full.shape[2] // 64
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    for i in (0, 37, 63):
        want = _records(oracle.infer({name: x[i:i + 1]})[out][0, 0])
        got = _records(full[0, 0, i * keep:(i + 1) * keep])
        assert got.shape == want.shape, (i, got.shape, want.shape)
        assert np.array_equal(got[:, 0:2], want[:, 0:2]), i
        ok, msg = close(got[:, 2:], want[:, 2:], rtol=1e-4, atol=1e-5)
        assert ok, (i, msg)


# ---- f16x2 contraction on badly scaled data -------------------------------------------------------------------------

def _loguniform(rng, shape, lo, hi):
    mag = np.exp(rng.uniform(np.log(lo), np.log(hi), shape))
    return (mag * rng.choice([-1.0, 1.0], shape)).astype(np.float32)


@pytest.mark.parametrize('cin,cout,k', [(64, 96, 1), (512, 64, 1), (32, 128, 3)])
def test_f16x2_wide_dynamic_range(cin, cout, k):
    """The FP16 hi/lo split on data FP16 alone could not hold: magnitudes log-uniform over 1e-8 .. 6e4 (FP16 hi parts
    are subnormal below 6e-5, zero below 6e-8) with mixed signs inside every K run, plus exactly cancelling pairs.
    Reference = float64.  Bound, relative to the size of the terms (not of the possibly cancelled sum):
        |y - ref| <= 2^-19 * sum|a||w|  +  2^-34 * (sum|a| * max|w| + sum|w| * max|a|) / K
    2^-22 per dropped lo*lo term and per operand rounding, 2^-23 per truncating tensor-core accumulation step between two
    round-to-nearest promotions (16 steps), and the 2^-36 absolute floor of the scaled
    residual of operands below the FP16 normal range."""
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(cin + cout + k)
    n, h, w = 2, 12, 12
    x = _loguniform(rng, (n, cin, h, w), 1e-8, 6e4)
    wt = _loguniform(rng, (cout, cin, k, k), 1e-8, 6e4)
    # cancelling pairs inside one K run: channel 2j+1 = -channel 2j with equal weights for a quarter of the filters
    x[:, 1:cin // 2:2] = -x[:, 0:cin // 2 - 1:2]
    wt[:cout // 4, 1:cin // 2:2] = wt[:cout // 4, 0:cin // 2 - 1:2]
    pad = k // 2
    y = kernels.conv2d(kernels.upload(x), kernels.upload(wt), (1, 1), (pad, pad), (h, w), math=_cabi.MATH_F16X2)
    kernels.status_reset()
    got = np.asarray(y).astype(np.float64)
    assert np.all(np.isfinite(got))
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    ref = np.zeros((n, cout, h, w))
    mag = np.zeros((n, cout, h, w))
    suma = np.zeros((n, 1, h, w))
    w64 = wt.astype(np.float64)
    for ky in range(k):
        for kx in range(k):
            patch = xp[:, :, ky:ky + h, kx:kx + w]
            ref += np.einsum('nchw,oc->nohw', patch, w64[:, :, ky, kx])
            mag += np.einsum('nchw,oc->nohw', np.abs(patch), np.abs(w64[:, :, ky, kx]))
            suma += np.abs(patch).sum(axis=1, keepdims=True)
    K = cin * k * k
    sumw = np.abs(w64).sum(axis=(1, 2, 3)).reshape(1, cout, 1, 1)
    bound = 2.0 ** -19 * mag + 2.0 ** -34 * (suma * np.abs(w64).max() + sumw * np.abs(x).max()) / K
    err = np.abs(got - ref)
    worst = float((err / bound).max())
    assert worst <= 1.0, 'max err/bound = {:.3f}'.format(worst)
    # the same data through the FP32-range kernels (what the engine falls back to) obeys the same bound
    y2 = np.asarray(kernels.conv2d(kernels.upload(x), kernels.upload(wt), (1, 1), (pad, pad), (h, w), math=_cabi.MATH_SAFE)).astype(np.float64)
    assert float((np.abs(y2 - ref) / bound).max()) <= 1.0


# ---- INTEGRATION.md option B through raw ctypes -----------------------------------------------------------------------

@pytest.mark.parametrize('shape,kernel,strides,pb,pe,rounding', [
    ((2, 64, 28, 28), '3,3', '1,1', '1,1', '1,1', 'ceil'),
    ((1, 32, 57, 57), '3,3', '2,2', '0,0', '0,0', 'ceil'),
    ((3, 6, 11, 11), '2,2', '2,2', '0,0', '0,0', 'floor'),
])
def test_integration_option_b_maxpool_ctypes_only(tmp_path, shape, kernel, strides, pb, pe, rounding):
    """Runs tests/scripts/option_b_maxpool.py (the INTEGRATION.md snippet: ctypes + numpy only, no torch, none of this
    package's Python) in a fresh interpreter and compares with the oracle, bit for bit."""
    from oracle import ref_ops
    x = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
    xin, yout = tmp_path / 'x.npy', tmp_path / 'y.npy'
    np.save(xin, x)
    r = subprocess.run([sys.executable, os.path.join(REPO, 'tests', 'scripts', 'option_b_maxpool.py'), str(xin), str(yout),
                        strides, pb, pe, kernel, rounding, 'explicit'], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    data = {'kernel': kernel, 'strides': strides, 'pads_begin': pb, 'pads_end': pe, 'rounding_type': rounding, 'auto_pad': 'explicit'}
    want = ref_ops.maxpool(data, x)
    assert np.array_equal(np.load(yout), want)


# ---- split-K MatMul ----------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('m,k,n,act', [(1024, 6272, 512, 'relu'), (256, 1024, 1000, None), (1, 576, 64, 'relu'), (64, 6272, 512, None),
                                       (130, 1000, 10, None), (5, 4104, 36, 'relu')])
def test_matmul_split_k_vs_oracle(m, k, n, act):
    """Few output tiles -> several CTAs per tile over slices of K + a fixed-order reduction (bias and activation applied
    there).  Same tolerance class as the unsplit contraction, deterministic, and K tails that are not whole slots."""
    import ctypes as C
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(m + k + n)
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = (rng.standard_normal((n, k)) * np.sqrt(2.0 / k)).astype(np.float32)
    bias = (0.05 * rng.standard_normal((1, n))).astype(np.float32)
    need = C.c_size_t(0)
    _cabi.call('b200ov_matmul_workspace', m, n, k, C.byref(need))
    if (m, k, n) in ((1024, 6272, 512), (256, 1024, 1000), (64, 6272, 512)):
        assert need.value > 0, 'expected a split-K plan for this shape'
    da, db, dbias = kernels.upload(a), kernels.upload(b), kernels.upload(bias)
    fact = ('relu',) if act else None
    got = np.asarray(kernels.matmul(da, db, bias=dbias, act=fact))
    want = ref_ops.add(ref_ops.matmul({'transpose_a': 'false', 'transpose_b': 'true'}, a, b), bias)
    if act:
        want = ref_ops.relu(want)
    ok, msg = close(got, want)
    assert ok, msg
    for _ in range(3):
        assert np.array_equal(np.asarray(kernels.matmul(da, db, bias=dbias, act=fact)), got)
    # the unsplit kernel on the same data: same tolerance class (not bit-identical: different summation tree)
    plain = np.asarray(kernels.matmul(da, db, bias=dbias, act=fact, math=_cabi.MATH_SAFE))
    ok, msg = close(plain, want)
    assert ok, msg


# ---- liveness-planned arena ------------------------------------------------------------------------------------------

@pytest.mark.parametrize('model,batch', [('googlenet-v1', 16), ('ssd_mobilenet_v1_coco', 4), ('mnist_bn', 64)])
def test_buffer_reuse_is_bit_identical_and_smaller(model_dir, model, batch):
    """Released feature-map buffers are handed to later producers (planned CUDA-graph mode).  Results are bit-identical
    to the run that keeps every buffer, replay after replay, and the live working set shrinks."""
    from tools.synth_bin import synth_input
    x = synth_input(model, batch=batch, seed=12)
    x2 = synth_input(model, batch=batch, seed=13)
    net_a, exe_a = _load(model_dir, model, batch=batch, reuse_buffers=False)
    net_b, exe_b = _load(model_dir, model, batch=batch, reuse_buffers=True)
    name, out = net_a.inputs[0]['name'], net_a.outputs[0]['name']
    for xx in (x, x2, x):
        a = exe_a.infer({name: xx})[out]
        b = exe_b.infer({name: xx})[out]
        assert np.array_equal(a, b)
    keep_all, reuse = exe_a._arena.peak_bytes(), exe_b._arena.peak_bytes()
    assert exe_b._arena.reused > 0
    assert reuse < 0.6 * keep_all, (keep_all, reuse)
    # a second input signature (uint8) captures another graph over the same arena
    x8 = np.random.default_rng(1).integers(0, 256, x.shape, dtype=np.uint8)
    assert np.array_equal(exe_a.infer({name: x8})[out], exe_b.infer({name: x8})[out])
    assert np.array_equal(exe_b.infer({name: x})[out], exe_a.infer({name: x})[out])


# ---- TMA-staged MaxPool / depthwise tiles ------------------------------------------------------------------------------

def _maybe_slice(kernels, x_host, rng):
    """Upload an NCHW array as NHWC, half of the time as a channel slice of a wider (Concat-style) buffer."""
    n, c, h, w = x_host.shape
    if rng.integers(0, 2) == 0:
        return kernels.to_nhwc(kernels.upload(x_host))
    extra = 4 * int(rng.integers(1, 5))
    lead = 4 * int(rng.integers(0, 3))
    wide = np.full((n, lead + c + extra, h, w), np.float32(1e9))       # poison around the slice
    wide[:, lead:lead + c] = x_host
    buf = kernels.to_nhwc(kernels.upload(wide.astype(np.float32)))
    return kernels.channel_slice(buf, lead, c)


def test_tma_maxpool_tiles_vs_oracle():
    """pool_max_tma_kernel over the tile geometries the planner produces (several images per tile, row / column tiles,
    partial channel chunks, ceil-mode overhang beyond the padded tensor, zero padding that takes part in the max,
    channel-slice inputs and outputs): bit-exact against the oracle."""
    from oracle import ref_ops
    from pyopenvino_b200 import _cabi, common_def, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(21)
    cases = 0
    for hw in (7, 10, 14, 19, 28, 33, 56, 75, 112):
        for (k, s, p, rounding) in ((3, 1, 1, 'ceil'), (3, 2, 0, 'ceil'), (3, 2, 1, 'floor'), (2, 2, 0, 'floor'), (2, 1, 0, 'ceil'), (3, 1, 0, 'floor')):
            c = int(rng.choice([16, 36, 64, 96, 132]))
            n = int(rng.integers(1, 10))
            while n * c * hw * hw < (1 << 17):
                n += 1
            x = (rng.standard_normal((n, c, hw, hw)) * 2 - (1.5 if rng.integers(0, 2) else 0.0)).astype(np.float32)
            data = {'strides': '{0},{0}'.format(s), 'kernel': '{0},{0}'.format(k), 'pads_begin': '{0},{0}'.format(p),
                    'pads_end': '{0},{0}'.format(p), 'rounding_type': rounding, 'auto_pad': 'explicit'}
            want = ref_ops.maxpool(data, x)
            oh, ow = want.shape[2:]
            xd = _maybe_slice(kernels, x, rng)
            out = None
            if rng.integers(0, 2):
                out = kernels.channel_slice(kernels.new_nhwc(n, c + 8, oh, ow), 4, c)
            y = kernels.pool2d(xd, _cabi.POOL_MAX, (k, k), (s, s), (p, p), (p, p), (oh, ow), out=out)
            assert np.array_equal(np.asarray(y), want), (hw, k, s, p, rounding, n, c)
            cases += 1
    assert cases == 54


def test_tma_depthwise_tiles_vs_oracle():
    """dwconv3x3_tma_kernel: FP32 tolerance class against the oracle over the planner's tile geometries, both strides,
    asymmetric pads, bias + Clamp epilogue, channel-slice inputs."""
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(22)
    for hw in (10, 19, 38, 75, 150, 23):
        for s in (1, 2):
            c = int(rng.choice([32, 48, 128, 516]))
            n = int(rng.integers(1, 6))
            while n * c * hw * hw < (1 << 18):
                n += 1
            pb = (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
            pe = (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
            x = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
            w = (rng.standard_normal((c, 1, 1, 3, 3)) * 0.5).astype(np.float32)
            b = (0.1 * rng.standard_normal((1, c, 1, 1))).astype(np.float32)
            want = np.clip(ref_ops.groupconv_numpy(x, w, (s, s), pb, pe, 'explicit') + b, 0.0, 6.0)
            oh, ow = want.shape[2:]
            xd = _maybe_slice(kernels, x, rng)
            y = kernels.dwconv2d(xd, kernels.upload(w), (s, s), pb, (oh, ow), bias=kernels.upload(b), act=('clamp', 0.0, 6.0))
            ok, msg = close(np.asarray(y), want)
            assert ok, ((hw, s, c, n, pb, pe), msg)
