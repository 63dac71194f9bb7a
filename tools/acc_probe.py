import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
from pyopenvino_b200.inference_engine import IECore
from oracle import ref_ops
plugins = IECore().plugins.plugins
def node(x, w, s, p):
    data = {'strides': '%d, %d' % (s, s), 'dilations': '1, 1', 'pads_begin': '%d, %d' % (p, p), 'pads_end': '%d, %d' % (p, p), 'auto_pad': 'explicit'}
    return {'name': 'c', 'type': 'Convolution', 'data': data, 'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}}, 'output': {2: {'precision': 'FP32', 'dims': ()}}}
for (n, cin, hw, cout, k, relu_in) in [(2, 64, 56, 192, 3, True), (2, 512, 19, 512, 1, True), (1, 1024, 10, 1024, 1, True), (1, 512, 14, 128, 3, True), (2, 832, 7, 384, 1, False), (1, 160, 14, 320, 3, True), (8, 6272, 1, 512, 1, True)]:
    rng = np.random.default_rng(cin + k)
    x = rng.standard_normal((n, cin, hw, hw)).astype(np.float32)
    if relu_in: x = np.maximum(x, 0)          # post-ReLU activations: same-sign products pile up in the accumulator
    w = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    want = ref_ops.conv_special(x, w, (1, 1), (k // 2, k // 2), (k // 2, k // 2), 'explicit').astype(np.float64)
    x64 = x.astype(np.float64)
    for kt in ('f16x2', 'tf32x3', 'fp32'):
        got = np.asarray(plugins['Convolution'].compute(node(x, w, 1, k // 2), {0: x, 1: w}, kernel_type=kt)[2]).astype(np.float64)
        err = np.abs(got - want)
        tol = 1e-5 + 1e-4 * np.abs(want)
        print('K=%5d cout=%4d %-7s max|d| %.2e  max d/tol %.3f  mean d %.2e (signed %.2e)  rms(out) %.2f' % (cin * k * k, cout, kt, err.max(), (err / tol).max(), err.mean(), (got - want).mean(), want.std()))
