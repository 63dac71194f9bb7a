"""Developer tool: where CTA 0 of the f16x2 contraction kernel waits (needs a -DB200OV_F16_TRACE build:
B200OV_EXTRA_NVCC_FLAGS=-DB200OV_F16_TRACE python -m pyopenvino_b200.build --force).

    python tools/f16_trace.py CIN COUT K HW [BATCH [STRIDE]]
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyopenvino_b200 import _cabi, kernels, device as dev
from pyopenvino_b200.inference_engine import IECore
dev.init()
plugins = IECore().plugins.plugins
cin, cout, k, hw = [int(a) for a in sys.argv[1:5]]
B = int(sys.argv[5]) if len(sys.argv) > 5 else 64
stride = int(sys.argv[6]) if len(sys.argv) > 6 else 1
rng = np.random.default_rng(0)
x = rng.standard_normal((B, cin, hw, hw)).astype(np.float32)
w = rng.standard_normal((cout, cin, k, k)).astype(np.float32)
pad = k // 2
node = {'name': 't', 'type': 'Convolution', 'data': {'strides': '%d, %d' % (stride, stride), 'dilations': '1, 1', 'pads_begin': '%d, %d' % (pad, pad), 'pads_end': '%d, %d' % (pad, pad), 'auto_pad': 'explicit'},
        'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}}, 'output': {2: {'precision': 'FP32', 'dims': ()}}}
ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
import torch
for _ in range(3):
    plugins['Convolution'].compute(node, ins, kernel_type='f16x2')
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib = _cabi.load()
assert lib.b200ov_debug_f16_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(4, 8)
oh = (hw + 2 * pad - k) // stride + 1
units = k * ((k + 1) // 2) if cin <= 4 else k * k * ((cin + 7) // 8)
slots = (units + 3) // 4
tiles = ((B * oh * oh + 127) // 128) * ((cout + 127) // 128 if cout > 64 else 1)
per_cta = (tiles + 147) // 148
print('cin %d cout %d k %d hw %d batch %d: %d slots/tile, %d tiles, <= %d tiles per CTA, %d items per CTA' % (cin, cout, k, hw, B, slots, tiles, per_cta, per_cta * slots))
names = [('B loader', ['wait b_empty']), ('MMA', ['wait main drained (N=128)', 'descriptors + election before first MMA', 'wait B stage (named barrier)', 'wait A slot (named barrier)', 'tcgen05.fence::after', 'descriptors + MMAs + commits', 'gap between bursts (incl. waits)']),
         ('producer w4', ['wait a_empty', 'issue loads', 'wait tmem st', 'convert+store item0 (incl. waits)', 'convert+store item1 (incl. waits)']),
         ('epilogue w12', ['wait main_full', 'wait cross_full', 'wait tma store read'])]
for r, (role, keys) in enumerate(names):
    total = int(t[r, 7])
    print('%-14s total %9d clk  (%.0f clk per item)' % (role, total, total / max(1, per_cta * slots)))
    for i, kname in enumerate(keys):
        print('    %-40s %9d clk  %5.1f%%' % (kname, int(t[r, i]), 100.0 * t[r, i] / max(1, total)))

tl = (ctypes.c_longlong * (8 * 128))()
if hasattr(lib, 'b200ov_debug_f16_timeline') and lib.b200ov_debug_f16_timeline(tl) == 0:
    T = np.array(tl, dtype=np.int64).reshape(8, 128)
    t0 = T[0, 0]
    print('item | loads_issued slot_free published | gate_ready | mma_start mma_end   (cycles since the first load issue; item i uses A slot i%4)')
    for i in range(16, min(per_cta * slots, 56)):
        print('%4d | %8d %8d %8d | %8d | %8d %8d' % ((i,) + tuple(int(T[e, i] - t0) for e in range(6))))
    print('chunk | epilogue_got epilogue_released')
    for c in range(4, 14):
        print('%4d | %8d %8d' % (c, int(T[6, c] - t0), int(T[7, c] - t0)))
