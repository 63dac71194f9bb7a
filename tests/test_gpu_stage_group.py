"""GPU parity of two producer-side variants of the contraction kernel (conv_f16x2.cu; reference Convolution.py:57-87):

* staged 1x1 convolutions / MatMul (pixel tiles brought into shared memory by TMA instead of register gathers): must be
  BIT-identical to the register-gather kernel (same split, same K order) and within the FP32 tolerance of the oracle;
* pixel-group stems (C_in <= 4, horizontal stride 2: `group` adjacent output pixels are one GEMM row over a shared window):
  within the FP32 tolerance of the oracle for every group size, tap alignment and fallback case.
"""
import os

import numpy as np
import pytest

from conftest import close

pytestmark = pytest.mark.gpu


def _act_ref(want, act):
    if act == ('relu',):
        return np.maximum(want, 0)
    if act is not None:
        return np.clip(want, act[1], act[2])
    return want


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize('shape', [
    # n, cin, h, w, cout, act, sliced input / output
    (3, 192, 28, 28, 64, ('relu',), False),       # inception_3a/1x1
    (2, 480, 14, 14, 192, ('relu',), True),       # channel slice in, Concat slot out, 96-wide tiles
    (2, 528, 14, 14, 256, None, False),           # 128-wide tiles, two column tiles
    (5, 832, 7, 7, 384, ('relu',), False),        # M = 245: partial last row tile
    (1, 40, 5, 9, 24, None, False),               # C_in not a multiple of the 32-channel slot (zero fill), C_out % 32 != 0
    (2, 64, 56, 56, 64, ('clamp', -0.5, 0.75), False),      # conv2/3x3_reduce: many tiles per CTA
    (1, 512, 19, 19, 273, None, False),           # SSD class head: output pitch TMA cannot express
    (300, 1024, 1, 1, 1000, None, False),         # MatMul shape (loss3/classifier)
])
def test_staged_1x1_bit_identical_to_register_gather_and_vs_oracle(shape):
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, cin, h, w, cout, act, sliced = shape
    rng = np.random.default_rng(abs(hash(shape[:5])) % (1 << 31))
    x = rng.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)

    def run():
        if sliced:
            wide = np.full((n, cin + 16, h, w), np.float32(1e9))
            wide[:, 8:8 + cin] = x
            xd = kernels.channel_slice(kernels.to_nhwc(kernels.upload(wide)), 8, cin)
            out = kernels.channel_slice(kernels.new_nhwc(n, cout + 64, h, w), 32, cout)
        else:
            xd, out = kernels.to_nhwc(kernels.upload(x)), None
        return np.asarray(kernels.conv2d(xd, kernels.upload(wt), (1, 1), (0, 0), (h, w), bias=kernels.upload(b), act=act, out=out))

    staged = run()
    with _env(B200OV_F16_STAGE='0'):
        gathered = run()
    assert np.array_equal(staged, gathered), 'staged 1x1 differs from the register-gather kernel (max |d| = {})'.format(
        np.abs(staged - gathered).max())
    want = _act_ref(ref_ops.conv_special(x, wt, (1, 1), (0, 0), (0, 0), 'explicit') + b, act)
    ok, msg = close(staged, want)
    assert ok, msg


@pytest.mark.parametrize('model,batch', [('googlenet-v1', 8), ('ssd_mobilenet_v1_coco', 2), ('mnist_bn', 64)])
def test_models_staged_vs_register_gather_bit_identical(model_dir, model, batch):
    """Whole networks (sibling 1x1 convolutions as one staged contraction with (hi, lo) reduce outputs, SSD heads, MatMul):
    B200OV_F16_STAGE=0 must give the same bits."""
    from pyopenvino_b200.inference_engine import IECore
    from tools.synth_bin import synth_input
    ie = IECore()
    xml = os.path.join(model_dir, model + '.xml')
    x = synth_input(model, batch=batch, seed=7)
    outs = []
    for off in (None, '0'):
        with _env(**({} if off is None else {'B200OV_F16_STAGE': off})):
            net = ie.read_network(xml, xml[:-4] + '.bin')
            exe = ie.load_network(net, 'B200', batch_size=batch)
            name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
            outs.append(exe.infer({name: x})[out_name])
            outs.append(exe.infer({name: x})[out_name])          # the captured graph
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


@pytest.mark.parametrize('shape', [
    # n, cin, hw, cout, k, s, pads_begin, pads_end, act
    (2, 3, 224, 64, 7, 2, (3, 3), (3, 3), ('relu',)),     # GoogLeNet conv1: group 2, odd left padding
    (2, 3, 300, 32, 3, 2, (0, 0), (1, 1), ('clamp', 0.0, 6.0)),   # SSD conv0: ow = 150 -> group 3
    (1, 3, 64, 32, 3, 2, (1, 1), (1, 1), None),           # ow = 32 -> group 4, odd padding
    (3, 3, 32, 16, 5, 2, (2, 2), (2, 2), ('relu',)),      # C_out 16: group 4 with 64 columns
    (1, 4, 30, 8, 4, 2, (1, 1), (2, 2), None),            # even kernel, ow = 15 -> group 3
    (2, 2, 20, 40, 2, 2, (0, 0), (0, 0), None),           # C_out 40: group 2, 80 columns (one 96-wide tile)
    (2, 3, 38, 64, 7, 2, (3, 3), (3, 3), None),           # ow = 19: no group divides it -> per-pixel rows
    (2, 1, 28, 24, 3, 2, (1, 1), (1, 1), ('relu',)),      # one input channel
])
def test_pixel_group_stems_vs_oracle(shape):
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, cin, hw, cout, k, s, pb, pe, act = shape
    rng = np.random.default_rng(abs(hash(shape[:6])) % (1 << 31))
    x = rng.standard_normal((n, cin, hw, hw)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)
    oh = (hw + pb[0] + pe[0] - k) // s + 1
    want = _act_ref(ref_ops.conv_special(x, wt, (s, s), pb, pe, 'explicit') + b, act)
    # the stem path needs a pixel pitch of 4 values: kernels.to_nhwc gives it to 3-channel inputs; pad the others with zero planes
    xp, wp = x, wt
    if cin not in (3, 4):
        xp = np.zeros((n, 4, hw, hw), np.float32); xp[:, :cin] = x
        wp = np.zeros((cout, 4, k, k), np.float32); wp[:, :cin] = wt
    outs = {}
    for g in ('4', '3', '2', '1'):
        with _env(B200OV_F16_STEM_GROUP=g):
            xd = kernels.to_nhwc(kernels.upload(xp))
            got = np.asarray(kernels.conv2d(xd, kernels.upload(wp), (s, s), pb, (oh, oh), bias=kernels.upload(b), act=act))
        ok, msg = close(got, want)
        assert ok, ('group <= ' + g, msg)
        outs[g] = got
    # the grouped forms only reorder zero terms inside the K run: tiny differences at most
    assert np.abs(outs['4'] - outs['1']).max() <= 1e-5 + 1e-4 * np.abs(want).max()


@pytest.mark.parametrize('shapes,axis', [
    ([(3, 1083 * 7), (3, 600 * 7), (3, 150 * 7), (3, 54 * 7), (3, 24 * 7), (3, 6 * 7)], 1),     # SSD head Concat: one launch
    ([(2, 8, 16), (2, 12, 16), (2, 4, 16)], 1),                                                # vector path, inner dims folded
    ([(5, 3)] * 9, 1),                                                                         # more parts than one launch takes
    ([(4, 10), (4, 1)], 1),
])
def test_row_concat_one_launch_bit_exact(shapes, axis):
    """Concat.py:9-13 on device-resident non-NHWC parts: b200ov_concat_rows must equal np.concatenate bit for bit."""
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200.inference_engine import IECore
    plugins = IECore().plugins.plugins
    rng = np.random.default_rng(len(shapes))
    parts = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    node = {'name': 'cat', 'type': 'Concat', 'data': {'axis': str(axis)},
            'input': {i: {'precision': 'FP32', 'dims': p.shape} for i, p in enumerate(parts)},
            'output': {len(parts): {'precision': 'FP32', 'dims': ()}}}
    launches0 = _cabi.launch_count
    got = plugins['Concat'].compute(node, {i: kernels.upload(p) for i, p in enumerate(parts)})[len(parts)]
    if len(parts) <= _cabi.CONCAT_MAX_PARTS:
        assert _cabi.launch_count - launches0 == 1
    assert np.array_equal(np.asarray(got), np.concatenate(parts, axis=axis))


def test_detection_output_batch_wide_top1_pass_same_records():
    """b200ov_detection_output_ws (grid-wide top-1 pass + per-image CTAs) against the single-kernel entry point: same
    records bit for bit (DetectionOutput.py:162-300; the golden-vector tests pin the ws path to the reference)."""
    import ctypes as C
    from pyopenvino_b200 import _cabi, kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(11)
    n, priors, classes, keep = 5, 1917, 91, 100
    loc = (0.5 * rng.standard_normal((n, priors * 4))).astype(np.float32)
    conf = (1.0 / (1.0 + np.exp(-(rng.standard_normal((n, priors * classes)) * 2 - 3)))).astype(np.float32)
    conf[:, ::7] = conf[:, 3::7]                   # exact ties between classes and between priors
    cx, cy = rng.random(priors), rng.random(priors)
    w, h = 0.05 + 0.3 * rng.random(priors), 0.05 + 0.3 * rng.random(priors)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1).astype(np.float32)
    var = np.tile(np.float32([0.1, 0.1, 0.2, 0.2]), (priors, 1))
    prop = np.stack([boxes.reshape(-1), var.reshape(-1)])[None].astype(np.float32)
    ld, cd, pd = kernels.upload(loc), kernels.upload(conf), kernels.upload(prop)
    got = np.asarray(kernels.detection_output(ld, cd, pd, classes, keep, True, False, False, True, 0.3, 0.6))
    out = kernels.upload(np.zeros((n * keep, 7), np.float32))
    d = _cabi.DetectionDesc(n=n, num_priors=priors, num_classes=classes, keep_top_k=keep, code_center_size=1,
                            variance_encoded_in_target=0, clip_before_nms=0, clip_after_nms=1,
                            confidence_threshold=0.3, nms_threshold=0.6)
    _cabi.call('b200ov_detection_output', C.byref(d), C.c_void_p(ld.ptr), C.c_void_p(cd.ptr), C.c_void_p(pd.ptr), C.c_void_p(out.ptr),
               C.c_void_p(dev.stream()))
    want = np.asarray(out).reshape(got.shape)
    assert np.array_equal(got, want)
    assert (got[0, 0, :, 0] >= 0).sum() > n        # the case really produces detections


@pytest.mark.parametrize('shape', [
    # n, hw, cout, k, s, pads_begin, pads_end, act
    (64, 28, 32, 3, 1, (1, 1), (1, 1), ('relu',)),        # mnist_bn conv2d
    (3, 28, 32, 5, 1, (0, 0), (0, 0), None),              # mnist conv (5x5, valid)
    (2, 31, 64, 7, 2, (3, 3), (3, 3), ('clamp', -0.2, 0.4)),
    (1, 9, 4, 2, 1, (0, 0), (1, 1), None),
])
def test_c1_direct_stem_vs_oracle_and_tiled_kernel(shape):
    """C_in = 1 stems: the direct kernel against the oracle (Convolution.py:57-87) and against the tiled FFMA kernel."""
    from oracle import ref_ops
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    n, hw, cout, k, s, pb, pe, act = shape
    rng = np.random.default_rng(abs(hash(shape[:5])) % (1 << 31))
    x = rng.standard_normal((n, 1, hw, hw)).astype(np.float32)
    wt = (rng.standard_normal((cout, 1, k, k)) * np.sqrt(2.0 / (k * k))).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, cout, 1, 1))).astype(np.float32)
    oh = (hw + pb[0] + pe[0] - k) // s + 1
    want = _act_ref(ref_ops.conv_special(x, wt, (s, s), pb, pe, 'explicit') + b, act)

    def run():
        return np.asarray(kernels.conv2d(kernels.to_nhwc(kernels.upload(x)), kernels.upload(wt), (s, s), pb, (oh, oh),
                                         bias=kernels.upload(b), act=act))
    direct = run()
    with _env(B200OV_NO_C1_DIRECT='1'):
        tiled = run()
    ok, msg = close(direct, want)
    assert ok, msg
    ok, msg = close(direct, tiled)
    assert ok, msg


@pytest.mark.parametrize('k', [3, 5])
def test_c1_direct_stem_hl_output_feeds_contraction_bit_identically(k):
    """C_in = 1 stem -> 3x3 convolution: the stem's direct kernel writing the (hi, lo) operand form must give the consumer
    the same bits as the FP32 edge (the split is the one the consumer's producers would apply)."""
    from pyopenvino_b200 import kernels
    from pyopenvino_b200 import device as dev
    dev.init()
    rng = np.random.default_rng(k)
    n, hw, c1, c2 = 6, 28, 32, 64
    x = rng.standard_normal((n, 1, hw, hw)).astype(np.float32)
    w1 = (rng.standard_normal((c1, 1, k, k)) * np.sqrt(2.0 / (k * k))).astype(np.float32)
    b1 = (0.1 * rng.standard_normal((1, c1, 1, 1))).astype(np.float32)
    w2 = (rng.standard_normal((c2, c1, 3, 3)) * np.sqrt(2.0 / (9 * c1))).astype(np.float32)
    p = k // 2
    outs = []
    for hl in (True, False):
        y1 = kernels.conv2d(kernels.to_nhwc(kernels.upload(x)), kernels.upload(w1), (1, 1), (p, p), (hw, hw), bias=kernels.upload(b1),
                            act=('relu',), hl_out=hl)
        assert y1.st == ('hl' if hl else 'f32')
        outs.append(np.asarray(kernels.conv2d(y1, kernels.upload(w2), (1, 1), (1, 1), (hw, hw))))
    assert np.array_equal(outs[0], outs[1])
