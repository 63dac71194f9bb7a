"""Device-resident tensors handed from node to node ("HBM-resident handle").

A `DeviceArray` quacks like the ndarray the reference plugins exchange as far as the plugin
contract needs: `.shape` is the logical IR shape (NCHW for feature maps), `.dtype` a numpy dtype,
`np.asarray(x)` gives the host copy in logical layout.  Physically, 4-D float32 feature maps live
in HBM as NHWC with a channel pitch `ld` (so a producer can write into a channel slice of a Concat
buffer); everything else is stored row-major ("plain").

PyTorch is used for plumbing only: device memory (`torch.empty`), the current CUDA stream and, in
`distributed.py`, NCCL.  All arithmetic and layout changes go through libb200ov (ctypes).
"""
import os

import numpy as np
import torch

from . import _cabi

_initialized = False


def init(device=None):
    """Bind this process to one GPU (LOCAL_RANK under torchrun).  Raises when there is no CUDA device:
    there is no CPU fallback."""
    global _initialized
    if _initialized:
        return
    if not torch.cuda.is_available():
        raise _cabi.B200ovError('pyopenvino_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(device)
    _cabi.load()
    _cabi.call('b200ov_init', device)
    _initialized = True


def stream():
    """The stream every kernel is launched on: torch's current stream (so torch.cuda.Event timing and
    CUDA-graph capture see our launches)."""
    return torch.cuda.current_stream().cuda_stream


def synchronize():
    """Wait for everything queued on the launch stream."""
    torch.cuda.current_stream().synchronize()


class Arena:
    """Bump allocator over a few large torch buffers.  `reset()` rewinds it; an identical sequence of
    `alloc` calls then returns identical addresses, which is what lets one eager warm-up run size the
    arena and a second, captured run be replayed as a CUDA graph."""

    ALIGN = 64            # floats (256 B)
    BLOCK = 64 << 20      # floats per block (256 MB) unless a single request is larger

    def __init__(self):
        self.blocks = []
        self.block_idx = 0
        self.cursor = 0
        self.frozen = False
        self.high_water = 0

    def reset(self):
        self.block_idx = 0
        self.cursor = 0

    def alloc(self, nfloats):
        n = max(int(nfloats), 1)
        n = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        while True:
            if self.block_idx < len(self.blocks):
                blk = self.blocks[self.block_idx]
                if self.cursor + n <= blk.numel():
                    out = blk[self.cursor:self.cursor + n]
                    self.cursor += n
                    return out
                self.block_idx += 1
                self.cursor = 0
                continue
            if self.frozen:
                raise _cabi.B200ovError('arena grew during CUDA-graph capture/replay; the warm-up run did not cover this allocation')
            self.blocks.append(torch.empty(max(n, self.BLOCK), dtype=torch.float32, device='cuda'))

    def bytes(self):
        return sum(b.numel() for b in self.blocks) * 4


_arena = None


def set_arena(arena):
    """Route `alloc_f32` through `arena` (None = torch's caching allocator)."""
    global _arena
    _arena = arena


def alloc_f32(nfloats):
    if _arena is not None:
        return _arena.alloc(nfloats)
    return torch.empty(max(int(nfloats), 1), dtype=torch.float32, device='cuda')


class DeviceArray:
    """float32 tensor resident in HBM.

    layout 'plain': `t` holds prod(shape) floats, row-major in logical order.
    layout 'nhwc' : logical shape (N, C, H, W); `t` holds N*H*W*ld floats; this array's channels are
                    [c_off, c_off + C) of each pixel.
    """
    __slots__ = ('t', 'shape', 'layout', 'ld', 'c_off', 'cache')
    dtype = np.dtype(np.float32)

    def __init__(self, t, shape, layout='plain', ld=None, c_off=0):
        self.t = t
        self.shape = tuple(int(s) for s in shape)
        self.layout = layout
        self.ld = int(ld) if ld is not None else (self.shape[1] if layout == 'nhwc' else 0)
        self.c_off = int(c_off)
        self.cache = {}

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.c_off

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1

    @property
    def pixels(self):
        n, c, h, w = self.shape
        return n * h * w

    def is_dense(self):
        return self.layout == 'plain' or (self.ld == self.shape[1] and self.c_off == 0)

    # ---- host boundary ------------------------------------------------------------------------
    def numpy(self):
        """Host copy in logical layout (the D2H edge of the graph)."""
        from . import kernels
        src = kernels.to_plain(self) if self.layout == 'nhwc' else self
        host = torch.empty(src.size, dtype=torch.float32).pin_memory() if src.size else torch.empty(0)
        if src.size:
            host.copy_(src.t[:src.size], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return host.numpy().reshape(self.shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return 'DeviceArray(shape={}, layout={}, ld={}, c_off={})'.format(self.shape, self.layout, self.ld, self.c_off)


def is_device(x):
    return isinstance(x, DeviceArray)
