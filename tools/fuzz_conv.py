"""Randomised parity sweep of the Convolution plugin (default arithmetic) against the oracle:
random batch / channels / image size / kernel / stride / padding / fused bias + activation.

    python tools/fuzz_conv.py [--cases 150] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops  # noqa: E402
from pyopenvino_b200 import kernels  # noqa: E402
from pyopenvino_b200.inference_engine import IECore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--cases', type=int, default=150)
ap.add_argument('--seed', type=int, default=0)
args = ap.parse_args()
plugins = IECore().plugins.plugins
rng = np.random.default_rng(args.seed)
bad = 0
for case in range(args.cases):
    k = int(rng.choice([1, 1, 3, 3, 5, 7, 2, 4]))
    s = int(rng.choice([1, 1, 2]))
    cin = int(rng.choice([1, 2, 3, 4, 8, 16, 24, 32, 40, 64, 96, 128, 160, 192, 256, 480]))
    cout = int(rng.choice([1, 8, 10, 16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 208, 256, 288, 320, 384]))
    hw = int(rng.integers(max(k, 3), 40))
    if cin >= 256:
        hw = min(hw, 16)
    n = int(rng.integers(1, 5))
    pb = (int(rng.integers(0, k // 2 + 1)), int(rng.integers(0, k // 2 + 1)))
    pe = (int(rng.integers(0, k // 2 + 1)), int(rng.integers(0, k // 2 + 1)))
    if hw + pb[0] + pe[0] < k or hw + pb[1] + pe[1] < k:
        continue
    x = rng.standard_normal((n, cin, hw, hw)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
            'pads_end': '{}, {}'.format(*pe), 'auto_pad': 'explicit'}
    node = {'name': 'conv', 'type': 'Convolution', 'data': data,
            'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
            'output': {2: {'precision': 'FP32', 'dims': ()}}}
    want = ref_ops.conv_special(x, w, (s, s), pb, pe, 'explicit')
    fused = None
    if rng.random() < 0.5:
        b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
        want = want + b.reshape(1, -1, 1, 1)
        act = rng.choice(['none', 'relu', 'clamp'])
        fused = {'bias': kernels.upload(b.reshape(1, cout, 1, 1))}
        if act == 'relu':
            want = np.where(want < 0, 0, want)
            fused['act'] = ('relu',)
        elif act == 'clamp':
            want = np.clip(want, 0.0, 6.0)
            fused['act'] = ('clamp', 0.0, 6.0)
    ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
    got = np.asarray(plugins['Convolution'].compute(node, ins, kernel_type='numpy', fused=fused)[2])
    err = np.abs(got - want)
    tol = 1e-5 + 1e-4 * np.abs(want)
    ok = got.shape == want.shape and bool(np.all(err <= tol))
    if not ok:
        bad += 1
        print('MISMATCH n={} cin={} hw={} cout={} k={} s={} pb={} pe={} fused={}: max err/tol {:.3g}'.format(
            n, cin, hw, cout, k, s, pb, pe, None if fused is None else fused.get('act', 'bias'), float((err / tol).max())))
print('{} cases, {} mismatches'.format(args.cases, bad))
sys.exit(1 if bad else 0)
