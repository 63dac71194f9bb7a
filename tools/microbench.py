#!/usr/bin/env python
"""Per-op micro-benchmark sweep (BASELINE.json configs[4]): Convolution 3x3 / 1x1 / 5x5 / 7x7, depthwise 3x3,
MatMul, MaxPool, AvgPool, LRN at GoogLeNet / SSD-MobileNet layer shapes, device-resident inputs, CUDA-event
timing, reported as GB/s vs measured HBM peak and TFLOP/s vs measured dense bf16 peak.

    python tools/microbench.py [--batch 64] [--iters 20] [--only conv] [--math tf32x3] [--out profiles/x.json]

Between timed iterations a 256 MB buffer is written to flush the 126 MB L2.
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# (tag, kind, params) -- per-image shapes from SURVEY.md section 8(d)
LAYERS = [
    ('G conv1 7x7s2', 'conv', dict(cin=3, cout=64, k=7, s=2, pb=3, pe=3, hw=224)),
    ('G conv2/3x3_reduce', 'conv', dict(cin=64, cout=64, k=1, s=1, pb=0, pe=0, hw=56)),
    ('G conv2/3x3', 'conv', dict(cin=64, cout=192, k=3, s=1, pb=1, pe=1, hw=56)),
    ('G 3a/1x1', 'conv', dict(cin=192, cout=64, k=1, s=1, pb=0, pe=0, hw=28)),
    ('G 3b/3x3', 'conv', dict(cin=128, cout=192, k=3, s=1, pb=1, pe=1, hw=28)),
    ('G 3a/5x5', 'conv', dict(cin=16, cout=32, k=5, s=1, pb=2, pe=2, hw=28)),
    ('G 4e/1x1', 'conv', dict(cin=528, cout=256, k=1, s=1, pb=0, pe=0, hw=14)),
    ('G 4e/3x3', 'conv', dict(cin=160, cout=320, k=3, s=1, pb=1, pe=1, hw=14)),
    ('G 5b/1x1', 'conv', dict(cin=832, cout=384, k=1, s=1, pb=0, pe=0, hw=7)),
    ('G 5b/3x3', 'conv', dict(cin=192, cout=384, k=3, s=1, pb=1, pe=1, hw=7)),
    ('S conv0 3x3s2', 'conv', dict(cin=3, cout=32, k=3, s=2, pb=0, pe=1, hw=300)),
    ('S pw1', 'conv', dict(cin=32, cout=64, k=1, s=1, pb=0, pe=0, hw=150)),
    ('S pw3', 'conv', dict(cin=128, cout=128, k=1, s=1, pb=0, pe=0, hw=75)),
    ('S pw5', 'conv', dict(cin=256, cout=256, k=1, s=1, pb=0, pe=0, hw=38)),
    ('S pw7-11', 'conv', dict(cin=512, cout=512, k=1, s=1, pb=0, pe=0, hw=19)),
    ('S pw13', 'conv', dict(cin=1024, cout=1024, k=1, s=1, pb=0, pe=0, hw=10)),
    ('S cls0', 'conv', dict(cin=512, cout=273, k=1, s=1, pb=0, pe=0, hw=19)),
    ('S dw1 s1', 'dw', dict(c=32, s=1, pb=1, pe=1, hw=150)),
    ('S dw2 s2', 'dw', dict(c=64, s=2, pb=0, pe=1, hw=150)),
    ('S dw3 s1', 'dw', dict(c=128, s=1, pb=1, pe=1, hw=75)),
    ('S dw5 s1', 'dw', dict(c=256, s=1, pb=1, pe=1, hw=38)),
    ('S dw7-11 s1', 'dw', dict(c=512, s=1, pb=1, pe=1, hw=19)),
    ('S dw13 s1', 'dw', dict(c=1024, s=1, pb=1, pe=1, hw=10)),
    ('G pool1 3x3s2', 'maxpool', dict(c=64, k=3, s=2, p=0, hw=112)),
    ('G 3a/pool 3x3s1', 'maxpool', dict(c=192, k=3, s=1, p=1, hw=28)),
    ('G 5a/pool 3x3s1', 'maxpool', dict(c=832, k=3, s=1, p=1, hw=7)),
    # pool -> pool_proj pairs of the inception modules: fused (MaxPool inside the 1x1 convolution's A producers) / separate
    ('G 3a/pool+proj fused', 'poolconv', dict(cin=192, cout=32, hw=28)),
    ('G 3a/pool+proj sep', 'poolconv_sep', dict(cin=192, cout=32, hw=28)),
    ('G 3b/pool+proj fused', 'poolconv', dict(cin=256, cout=64, hw=28)),
    ('G 3b/pool+proj sep', 'poolconv_sep', dict(cin=256, cout=64, hw=28)),
    ('G 4a/pool+proj fused', 'poolconv', dict(cin=480, cout=64, hw=14)),
    ('G 4a/pool+proj sep', 'poolconv_sep', dict(cin=480, cout=64, hw=14)),
    ('G 4e/pool+proj fused', 'poolconv', dict(cin=528, cout=128, hw=14)),
    ('G 4e/pool+proj sep', 'poolconv_sep', dict(cin=528, cout=128, hw=14)),
    ('G 5a/pool+proj fused', 'poolconv', dict(cin=832, cout=128, hw=7)),
    ('G 5a/pool+proj sep', 'poolconv_sep', dict(cin=832, cout=128, hw=7)),
    ('G pool5 avg7', 'avgpool', dict(c=1024, k=7, s=1, p=0, hw=7)),
    ('G norm1 lrn', 'lrn', dict(c=64, hw=56)),
    ('G norm2 lrn', 'lrn', dict(c=192, hw=56)),
    ('G FC 1024x1000', 'matmul', dict(k=1024, n=1000)),
    ('bn FC1 6272x512', 'matmul', dict(k=6272, n=512)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--only', default=None)
    ap.add_argument('--math', default='numpy')
    ap.add_argument('--out', default=None)
    ap.add_argument('--no-flush', action='store_true')
    args = ap.parse_args()

    import torch
    from pyopenvino_b200 import device as dev, kernels
    from pyopenvino_b200.inference_engine import IECore
    dev.init()
    plugins = IECore().plugins.plugins
    peaks_path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    peaks = json.load(open(peaks_path)) if os.path.isfile(peaks_path) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
    rng = np.random.default_rng(0)
    flush = torch.empty(128 << 20, dtype=torch.float32, device='cuda')   # 512 MB: flushes L2 and gives the host a head start
    stream = torch.cuda.Stream()
    B = args.batch
    rows = []
    with torch.cuda.stream(stream):
        for tag, kind, q in LAYERS:
            if args.only and args.only not in kind and args.only not in tag:
                continue
            fp = lambda dims: {'precision': 'FP32', 'dims': tuple(dims)}
            if kind == 'conv':
                x = rng.standard_normal((B, q['cin'], q['hw'], q['hw'])).astype(np.float32)
                w = (rng.standard_normal((q['cout'], q['cin'], q['k'], q['k'])) * np.sqrt(2.0 / (q['cin'] * q['k'] ** 2))).astype(np.float32)
                node = {'name': tag, 'type': 'Convolution', 'data': {'strides': '{0}, {0}'.format(q['s']), 'dilations': '1, 1',
                        'pads_begin': '{0}, {0}'.format(q['pb']), 'pads_end': '{0}, {0}'.format(q['pe']), 'auto_pad': 'explicit'},
                        'input': {0: fp(x.shape), 1: fp(w.shape)}, 'output': {2: fp(())}}
                ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
                bias = kernels.upload(np.zeros((1, q['cout'], 1, 1), np.float32))
                fused = {'bias': bias, 'act': ('relu',)}
                typ = 'Convolution'
            elif kind in ('poolconv', 'poolconv_sep'):
                x = np.maximum(rng.standard_normal((B, q['cin'], q['hw'], q['hw'])), 0).astype(np.float32)
                w = (rng.standard_normal((q['cout'], q['cin'], 1, 1)) * np.sqrt(2.0 / q['cin'])).astype(np.float32)
                node = {'name': tag, 'type': 'Convolution', 'data': {'strides': '1, 1', 'dilations': '1, 1', 'pads_begin': '0, 0',
                        'pads_end': '0, 0', 'auto_pad': 'explicit'}, 'input': {0: fp(x.shape), 1: fp(w.shape)}, 'output': {2: fp(())}}
                pnode = {'name': tag + ' pool', 'type': 'MaxPool', 'data': {'strides': '1, 1', 'kernel': '3, 3', 'pads_begin': '1, 1',
                         'pads_end': '1, 1', 'rounding_type': 'ceil', 'auto_pad': 'explicit'}, 'input': {0: fp(x.shape)}, 'output': {1: fp(())}}
                ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
                fused = {'bias': kernels.upload(np.zeros((1, q['cout'], 1, 1), np.float32)), 'act': ('relu',)}
                if kind == 'poolconv':
                    fused['pre_pool'] = 0
                typ = 'Convolution'
            elif kind == 'dw':
                x = rng.standard_normal((B, q['c'], q['hw'], q['hw'])).astype(np.float32)
                w = rng.standard_normal((q['c'], 1, 1, 3, 3)).astype(np.float32)
                node = {'name': tag, 'type': 'GroupConvolution', 'data': {'strides': '{0}, {0}'.format(q['s']), 'dilations': '1, 1',
                        'pads_begin': '{0}, {0}'.format(q['pb']), 'pads_end': '{0}, {0}'.format(q['pe']), 'auto_pad': 'explicit'},
                        'input': {0: fp(x.shape), 1: fp(w.shape)}, 'output': {2: fp(())}}
                ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
                fused = {'bias': kernels.upload(np.zeros((1, q['c'], 1, 1), np.float32)), 'act': ('clamp', 0.0, 6.0)}
                typ = 'GroupConvolution'
            elif kind in ('maxpool', 'avgpool'):
                x = rng.standard_normal((B, q['c'], q['hw'], q['hw'])).astype(np.float32)
                node = {'name': tag, 'type': 'MaxPool', 'data': {'strides': '{0}, {0}'.format(q['s']), 'kernel': '{0}, {0}'.format(q['k']),
                        'pads_begin': '{0}, {0}'.format(q['p']), 'pads_end': '{0}, {0}'.format(q['p']), 'rounding_type': 'ceil',
                        'auto_pad': 'explicit'}, 'input': {0: fp(x.shape)}, 'output': {1: fp(())}}
                ins = {0: kernels.to_nhwc(kernels.upload(x))}
                fused = None
                typ = 'MaxPool' if kind == 'maxpool' else 'AvgPool'
            elif kind == 'lrn':
                x = rng.standard_normal((B, q['c'], q['hw'], q['hw'])).astype(np.float32)
                node = {'name': tag, 'type': 'LRN', 'data': {'alpha': '9.9999997473787516e-05', 'beta': '0.75', 'bias': '1', 'size': '5'},
                        'input': {0: fp(x.shape), 1: {'precision': 'I64', 'dims': (1,)}}, 'output': {2: fp(())}}
                ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: np.array([1], dtype=np.int64)}
                fused = None
                typ = 'LRN'
            else:
                m = max(B, 1)
                a = rng.standard_normal((m, q['k'])).astype(np.float32)
                b = (rng.standard_normal((q['n'], q['k'])) * np.sqrt(2.0 / q['k'])).astype(np.float32)
                node = {'name': tag, 'type': 'MatMul', 'data': {'transpose_a': 'false', 'transpose_b': 'true'},
                        'input': {0: fp(a.shape), 1: fp(b.shape)}, 'output': {2: fp(())}}
                ins = {0: kernels.upload(a), 1: kernels.upload(b)}
                fused = {'bias': kernels.upload(np.zeros((1, q['n']), np.float32)), 'act': ('relu',)}
                typ = 'MatMul'
            call = (lambda: plugins[typ].compute(node, ins, kernel_type=args.math, fused=fused)) if fused is not None else \
                (lambda: plugins[typ].compute(node, ins, kernel_type=args.math))
            if kind == 'poolconv_sep':
                def call(node=node, pnode=pnode, ins=ins, fused=fused):
                    pooled = plugins['MaxPool'].compute(pnode, {0: ins[0]}, kernel_type=args.math)[1]
                    return plugins['Convolution'].compute(node, {0: pooled, 1: ins[1]}, kernel_type=args.math, fused=fused)
            try:
                y = next(iter(call().values()))
            except Exception as e:          # e.g. tcgen05 modes on C_in = 3 stems
                rows.append({'layer': tag, 'kind': kind, 'error': str(e)[:80]})
                continue
            in_elems = sum(int(np.prod(v.shape)) for v in ins.values() if hasattr(v, 'layout'))
            nbytes = 4 * (in_elems + y.size)
            if kind == 'conv':
                flops = 2 * y.size * q['cin'] * q['k'] ** 2
            elif kind in ('poolconv', 'poolconv_sep'):
                flops = 2 * y.size * q['cin']
            elif kind == 'dw':
                flops = 2 * y.size * 9
            elif kind == 'matmul':
                flops = 2 * y.size * q['k']
            else:
                flops = 0
            for _ in range(3):
                call()
            times = []
            for _ in range(args.iters):
                if not args.no_flush:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                call()
                e1.record()
                e1.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            row = {'layer': tag, 'kind': kind, 'batch': B, 'ms': ms, 'mbytes': nbytes / 1e6, 'gflop': flops / 1e9,
                   'gbs': nbytes / ms / 1e6, 'tflops': flops / ms / 1e9, 'frac_hbm': nbytes / ms / 1e6 / peaks['hbm_gbs'],
                   'frac_bf16_peak': flops / ms / 1e9 / peaks['bf16_tflops']}
            rows.append(row)
            print('{:22s} {:8s} {:8.3f} ms {:9.1f} MB {:8.1f} GF {:8.0f} GB/s ({:4.0%} hbm) {:7.1f} TF/s'.format(
                tag, kind, ms, row['mbytes'], row['gflop'], row['gbs'], row['frac_hbm'], row['tflops']), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({'batch': B, 'math': args.math, 'peaks': peaks, 'rows': rows}, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
