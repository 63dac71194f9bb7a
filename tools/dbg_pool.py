import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_ops
from pyopenvino_b200 import _cabi, kernels, device as dev
dev.init()
POOL_DATA = {'strides': '1,1', 'kernel': '3,3', 'pads_begin': '1,1', 'pads_end': '1,1', 'rounding_type': 'ceil', 'auto_pad': 'explicit'}
for (n, cin, h, w, cout) in [(2, 480, 14, 14, 64), (5, 832, 7, 7, 128), (2, 528, 14, 14, 128), (3, 192, 28, 28, 32)]:
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((n, cin, h, w)) - 0.8).astype(np.float32)
    # identity-like weights: output channel j = pooled channel j (so errors point at (pixel, channel))
    wt = np.zeros((cout, cin, 1, 1), np.float32)
    for j in range(cout):
        wt[j, j, 0, 0] = 1.0
    xd = kernels.to_nhwc(kernels.upload(x))
    wd = kernels.upload(wt)
    want = ref_ops.maxpool(POOL_DATA, x)[:, :cout]
    for rep in range(4):
        y = np.asarray(kernels.conv2d(xd, wd, (1, 1), (0, 0), (h, w), pre_pool=True))
        bad = np.argwhere(np.abs(y - want) > 1e-3)
        print((n, cin, h, w, cout), 'rep', rep, 'bad', len(bad), 'of', y.size)
        if len(bad):
            imgs = sorted(set(bad[:, 0].tolist())); chans = sorted(set(bad[:, 1].tolist()))
            pix = sorted(set((b[0] * h * w + b[2] * w + b[3]) for b in bad.tolist()))
            print('  images', imgs, 'channels', chans[:40], 'n_pix', len(pix), 'pix', pix[:48])
            b = bad[0]; print('  first', b, y[tuple(b)], want[tuple(b)])
