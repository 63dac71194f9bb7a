"""Randomised parity sweep of the bandwidth-bound plugins against the oracle: MaxPool (bit-exact), depthwise
GroupConvolution ('exact' bit-exact, default within the FP32 tolerance), LRN (tolerance).

    python tools/fuzz_ops.py [--cases 300] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops  # noqa: E402
from pyopenvino_b200.inference_engine import IECore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--cases', type=int, default=300)
ap.add_argument('--seed', type=int, default=0)
args = ap.parse_args()
plugins = IECore().plugins.plugins
rng = np.random.default_rng(args.seed)
bad = 0


def report(what, desc, detail):
    global bad
    bad += 1
    print('MISMATCH {} {}: {}'.format(what, desc, detail))


def close(got, want):
    err = np.abs(got - want)
    tol = 1e-5 + 1e-4 * np.abs(want)
    return got.shape == want.shape and bool(np.all(err <= tol)), float((err / tol).max()) if got.shape == want.shape else -1.0


for case in range(args.cases):
    kind = case % 3
    n = int(rng.integers(1, 4))
    c = int(rng.choice([1, 3, 4, 8, 12, 16, 32, 48, 64, 96, 128, 192, 256, 512]))
    hw = int(rng.integers(4, 34))
    x = (rng.standard_normal((n, c, hw, hw)) * 2).astype(np.float32)
    if kind == 0:
        k = int(rng.choice([2, 3, 3, 5]))
        s = int(rng.choice([1, 2, 2, 3]))
        p = int(rng.integers(0, k // 2 + 1))
        rounding = str(rng.choice(['ceil', 'floor']))
        if hw + 2 * p < k:
            continue
        data = {'strides': '{}, {}'.format(s, s), 'kernel': '{}, {}'.format(k, k), 'pads_begin': '{}, {}'.format(p, p),
                'pads_end': '{}, {}'.format(p, p), 'rounding_type': rounding, 'auto_pad': 'explicit'}
        node = {'name': 'pool', 'type': 'MaxPool', 'data': data, 'input': {0: {'precision': 'FP32', 'dims': x.shape}},
                'output': {1: {'precision': 'FP32', 'dims': ()}}}
        try:
            want = ref_ops.maxpool(data, x)
        except Exception:
            continue
        got = plugins['MaxPool'].compute(node, {0: x}, kernel_type='numpy')[1]
        if not np.array_equal(got, want):
            report('MaxPool', (n, c, hw, k, s, p, rounding), 'not bit-exact')
    elif kind == 1:
        s = int(rng.choice([1, 2]))
        pb = (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
        pe = (int(rng.integers(0, 2)), int(rng.integers(0, 2)))
        w = (rng.standard_normal((c, 1, 1, 3, 3)) * 0.5).astype(np.float32)
        data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
                'pads_end': '{}, {}'.format(*pe), 'auto_pad': 'explicit'}
        node = {'name': 'dw', 'type': 'GroupConvolution', 'data': data,
                'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
                'output': {2: {'precision': 'FP32', 'dims': ()}}}
        want = ref_ops.groupconv_numpy(x, w, (s, s), pb, pe, 'explicit')
        exact = plugins['GroupConvolution'].compute(node, {0: x, 1: w}, kernel_type='exact')[2]
        if not np.array_equal(exact, want):
            report('GroupConvolution exact', (n, c, hw, s, pb, pe), 'not bit-exact')
        ok, worst = close(plugins['GroupConvolution'].compute(node, {0: x, 1: w}, kernel_type='numpy')[2], want)
        if not ok:
            report('GroupConvolution', (n, c, hw, s, pb, pe), 'max err/tol {:.3g}'.format(worst))
    else:
        size = int(rng.choice([3, 5, 5, 7, 9]))
        data = {'alpha': str(float(rng.choice([1e-4, 9.9999997e-05, 5e-4]))), 'beta': '0.75', 'bias': '1.0', 'size': str(size)}
        node = {'name': 'norm', 'type': 'LRN', 'data': data, 'input': {0: {'precision': 'FP32', 'dims': x.shape}},
                'output': {2: {'precision': 'FP32', 'dims': ()}}}
        ok, worst = close(plugins['LRN'].compute(node, {0: x}, kernel_type='numpy')[2], ref_ops.lrn(data, x))
        if not ok:
            report('LRN', (n, c, hw, size), 'max err/tol {:.3g}'.format(worst))
print('{} cases, {} mismatches'.format(args.cases, bad))
sys.exit(1 if bad else 0)
