"""Developer tool: dump the tcgen05 kernel's per-role pipeline timestamps (needs a -DB200OV_TC_TRACE build)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyopenvino_b200 import _cabi, kernels, device as dev
from pyopenvino_b200.inference_engine import IECore
dev.init()
plugins = IECore().plugins.plugins
B, cin, cout, k, hw = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 192, int(sys.argv[3]) if len(sys.argv) > 3 else 3, int(sys.argv[4]) if len(sys.argv) > 4 else 28
rng = np.random.default_rng(0)
x = rng.standard_normal((B, cin, hw, hw)).astype(np.float32)
w = rng.standard_normal((cout, cin, k, k)).astype(np.float32)
pad = k // 2
node = {'name': 't', 'type': 'Convolution', 'data': {'strides': '1, 1', 'dilations': '1, 1', 'pads_begin': '%d, %d' % (pad, pad), 'pads_end': '%d, %d' % (pad, pad), 'auto_pad': 'explicit'},
        'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}}, 'output': {2: {'precision': 'FP32', 'dims': ()}}}
ins = {0: kernels.to_nhwc(kernels.upload(x)), 1: kernels.upload(w)}
for _ in range(3):
    plugins['Convolution'].compute(node, ins, kernel_type='tf32x3')
import torch
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (8 * 128))()
lib = _cabi.load()
assert lib.b200ov_debug_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(8, 128)
t0 = t[7, 0]
nkb = k * k * ((cin + 31) // 32)
print('setup %d  epilogue_start %d  end %d  (k-blocks %d)' % (t[7, 1] - t0, t[7, 2] - t0, t[7, 3] - t0, nkb))
print('kb | prod_wait_done prod_issued | conv_got_A conv_done promo_done | mma_ready mma_issued')
for kb in range(min(nkb, 40)):
    print('%3d | %7d %7d | %7d %7d %7d | %7d %7d' % ((kb,) + tuple(int(t[e, kb] - t0) for e in range(7))))

print('kb | mma_loop_top after_acc_free after_conv | ready')
for kb in range(min(nkb, 40)):
    print('%3d | %7d %7d %7d | %7d' % (kb, int(t[0, 64 + kb] - t0), int(t[1, 64 + kb] - t0), int(t[2, 64 + kb] - t0), int(t[5, kb] - t0)))
