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
    global _numa_cpus
    _numa_cpus = _gpu_local_cpus(device)
    if os.environ.get('B200OV_DEBUG_NUMA'):
        print('[b200ov] GPU {}: staging buffers bound to CPUs {}'.format(device, sorted(_numa_cpus) if _numa_cpus else None), flush=True)
    _initialized = True


_numa_cpus = None      # CPUs of the NUMA node this process's GPU hangs off (None: unknown / binding disabled)


def _gpu_local_cpus(device):
    """CPU set of the NUMA node the GPU is attached to (sysfs `local_cpulist` of its PCI function), or None."""
    if os.environ.get('B200OV_NO_NUMA_BIND'):
        return None
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = '{:04x}:{:02x}:{:02x}.0'.format(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/{}/local_cpulist'.format(bdf)) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(','):
            if not part:
                continue
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus if cpus and cpus != allowed else None
    except Exception:                       # no sysfs entry, old torch, container without the attribute: leave placement alone
        return None


def bind_host_thread(local_rank, local_world):
    """One process per GPU on one host: give each rank's python thread its own slice of the host CPUs (the GPU-local
    NUMA node's when sysfs exposes it, else an even split of the allowed set), so eight launch loops and their pinned
    staging copies do not migrate over each other.  Returns the CPU set or None when nothing was changed."""
    if os.environ.get('B200OV_NO_NUMA_BIND'):
        return None
    try:
        allowed = sorted(os.sched_getaffinity(0))
        local = _gpu_local_cpus(local_rank) if torch.cuda.is_available() else None
        pool = sorted(local) if local else allowed
        share = max(1, len(pool) // max(1, local_world))
        if local:
            # ranks whose GPUs share a NUMA node share its CPUs: split by position among all ranks (upper bound on sharers)
            start = (local_rank * share) % len(pool)
        else:
            start = (local_rank * share) % len(pool)
        mine = set(pool[start:start + share]) or set(pool)
        os.sched_setaffinity(0, mine)
        return mine
    except (OSError, AttributeError):
        return None


def pinned_empty(n, dtype=torch.float32, zero=False):
    """Page-locked host buffer for the H2D / D2H edges, allocated on the NUMA node of this process's GPU: the
    allocating thread is confined to the GPU-local CPUs while the pages are faulted in (first touch), then gets its
    affinity back.  With one process per GPU and default placement the staging buffers of some ranks land on the far
    socket and their copies cross the inter-socket link, which is what limited the 2- and 8-GPU end-to-end numbers."""
    global _numa_cpus
    prev = None
    if _numa_cpus:
        try:
            prev = os.sched_getaffinity(0)
            os.sched_setaffinity(0, _numa_cpus)
        except OSError:
            prev = None
    try:
        t = (torch.zeros if zero else torch.empty)(max(int(n), 1), dtype=dtype).pin_memory()
    finally:
        if prev is not None:
            os.sched_setaffinity(0, prev)
    return t


def stream():
    """The stream every kernel is launched on: torch's current stream (so torch.cuda.Event timing and
    CUDA-graph capture see our launches)."""
    return torch.cuda.current_stream().cuda_stream


def synchronize():
    """Wait for everything queued on the launch stream."""
    torch.cuda.current_stream().synchronize()


class Arena:
    """Per-inference buffer allocator over a few large torch buffers.

    `reset()` rewinds it; an identical sequence of `alloc` / `release` calls then returns identical addresses, which is
    what lets one eager warm-up run size the arena and a second, captured run be replayed as a CUDA graph.
    `release(t)` hands a buffer back (the executor calls it when the last consumer of a feature map has been queued:
    liveness-planned reuse, SURVEY.md 8(b) "ownership"); later requests take the best-fitting free chunk before the
    bump pointer moves.  Kernels run in stream order, so handing a chunk to a later producer is safe."""

    ALIGN = 64            # floats (256 B)
    BLOCK = 64 << 20      # floats per block (256 MB) unless a single request is larger

    def __init__(self):
        self.blocks = []
        self.frozen = False
        self.reset()
        self.high_water = 0          # floats in use at the peak of the last pass (live chunks, not block capacity)

    def reset(self):
        self.block_idx = 0
        self.cursor = 0
        self.free = []               # [(block, offset, size)] sorted by (block, offset)
        self.live = {}               # data_ptr -> (block, offset, size)
        self.in_use = 0
        self.peak = 0
        self.allocs = 0
        self.reused = 0

    def _take(self, blk, off, n):
        out = self.blocks[blk][off:off + n]
        self.live[out.data_ptr()] = (blk, off, n)
        self.in_use += n
        self.allocs += 1
        if self.in_use > self.peak:
            self.peak = self.in_use
            self.high_water = max(self.high_water, self.peak)
        return out

    def alloc(self, nfloats):
        n = max(int(nfloats), 1)
        n = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        best = -1
        for i, (blk, off, size) in enumerate(self.free):            # best fit among released chunks
            if size >= n and (best < 0 or size < self.free[best][2]):
                best = i
        if best >= 0:
            blk, off, size = self.free.pop(best)
            if size > n:
                self.free.append((blk, off + n, size - n))
                self.free.sort()
            self.reused += 1
            return self._take(blk, off, n)
        while True:
            if self.block_idx < len(self.blocks):
                cap = self.blocks[self.block_idx].numel()
                if self.cursor + n <= cap:
                    off = self.cursor
                    self.cursor += n
                    return self._take(self.block_idx, off, n)
                if self.cursor < cap:                               # the tail of this block stays usable for small requests
                    self.free.append((self.block_idx, self.cursor, cap - self.cursor))
                    self.free.sort()
                self.block_idx += 1
                self.cursor = 0
                continue
            if self.frozen:
                raise _cabi.B200ovError('arena grew during CUDA-graph capture/replay; the warm-up run did not cover this allocation')
            self.blocks.append(torch.empty(max(n, self.BLOCK), dtype=torch.float32, device='cuda'))

    def owns(self, t):
        return t.data_ptr() in self.live

    def release(self, t):
        """Give the chunk that starts at `t.data_ptr()` back; neighbouring free chunks are merged."""
        chunk = self.live.pop(t.data_ptr(), None)
        if chunk is None:
            return False
        self.in_use -= chunk[2]
        self.free.append(chunk)
        self.free.sort()
        merged = []
        for c in self.free:
            if merged and merged[-1][0] == c[0] and merged[-1][1] + merged[-1][2] == c[1]:
                merged[-1] = (c[0], merged[-1][1], merged[-1][2] + c[2])
            else:
                merged.append(c)
        self.free = merged
        return True

    def bytes(self):
        """Device memory held by the arena (block capacity)."""
        return sum(b.numel() for b in self.blocks) * 4

    def peak_bytes(self):
        """Bytes simultaneously live at the peak of the last pass (the working set)."""
        return self.peak * 4


_arena = None


def set_arena(arena):
    """Route `alloc_f32` through `arena` (None = torch's caching allocator)."""
    global _arena
    _arena = arena


def current_arena():
    return _arena


def alloc_f32(nfloats):
    if _arena is not None:
        return _arena.alloc(nfloats)
    return torch.empty(max(int(nfloats), 1), dtype=torch.float32, device='cuda')


class DeviceArray:
    """FP32 tensor of the graph, resident in HBM.

    layout 'plain': `t` holds prod(shape) floats, row-major in logical order.
    layout 'nhwc' : logical shape (N, C, H, W); `t` holds N*H*W*ld elements; this array's channels are
                    [c_off, c_off + C) of each pixel.
    `st` is the STORAGE type of an 'nhwc' feature map: 'f32' (default, the reference's precision) or 'f16' (opt-in
    storage mode, `load_network(..., storage='f16')`: half the bytes, every kernel still computes in FP32).  The logical
    'hl' marks a network input the layout kernel already split into the contraction's FP16 (hi, lo) pairs (same 4 bytes
    per value; only the stem Convolution reads it).  The logical
    `dtype` the plugin contract validates stays float32 either way; `t` is always a float32 torch tensor used as raw
    storage (an 'f16' map of n elements occupies ceil(n / 2) of its floats).
    """
    __slots__ = ('t', 'shape', 'layout', 'ld', 'c_off', 'cache', 'st')
    dtype = np.dtype(np.float32)

    def __init__(self, t, shape, layout='plain', ld=None, c_off=0, st='f32'):
        self.t = t
        self.shape = tuple(int(s) for s in shape)
        self.layout = layout
        self.ld = int(ld) if ld is not None else (self.shape[1] if layout == 'nhwc' else 0)
        self.c_off = int(c_off)
        self.cache = {}
        self.st = st
        assert st == 'f32' or layout == 'nhwc'

    @property
    def esize(self):
        return 2 if self.st == 'f16' else 4

    @property
    def code(self):
        return {'f32': _cabi.DT_F32, 'f16': _cabi.DT_F16, 'hl': _cabi.DT_HL}[self.st]

    @property
    def ptr(self):
        return self.t.data_ptr() + self.esize * self.c_off

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
        if self.st == 'hl':
            # debug view of a contraction-operand tensor: c = hi + 2^-11 lo (22 significant bits), never on the hot path
            n, c, h, w = self.shape
            words = self.t[:n * h * w * self.ld].view(torch.int32).view(-1, 4)
            hi = words[:, :2].contiguous().view(torch.float16).float()
            lo = words[:, 2:].contiguous().view(torch.float16).float()
            vals = (hi + lo * (1.0 / 2048.0)).view(n * h * w, self.ld)[:, self.c_off:self.c_off + c]
            return vals.reshape(n, h, w, c).permute(0, 3, 1, 2).contiguous().cpu().numpy()
        src = kernels.to_plain(self) if self.layout == 'nhwc' else self
        host = pinned_empty(src.size) if src.size else torch.empty(0)
        if src.size:
            host.copy_(src.t[:src.size], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return host.numpy().reshape(self.shape).copy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return 'DeviceArray(shape={}, layout={}, ld={}, c_off={}, st={})'.format(self.shape, self.layout, self.ld, self.c_off, self.st)


class RawInput:
    """A host input staged in HBM in its native element type (uint8 / int8 / float16 / float32), logical IR shape,
    row-major.  Only the Parameter plugin consumes it: the widening to float32 happens inside the layout kernel
    (`b200ov_input_to_nhwc`), i.e. after the PCIe copy, and is exact (`Parameter.py:13` does the same cast on the host)."""
    __slots__ = ('t', 'shape', 'np_dtype')
    CODES = {np.dtype(np.float32): _cabi.DT_F32, np.dtype(np.float16): _cabi.DT_F16, np.dtype(np.uint8): _cabi.DT_U8,
             np.dtype(np.int8): _cabi.DT_I8}
    TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float16): torch.float16, np.dtype(np.uint8): torch.uint8,
             np.dtype(np.int8): torch.int8}

    def __init__(self, t, shape, np_dtype):
        self.t = t
        self.shape = tuple(int(s) for s in shape)
        self.np_dtype = np.dtype(np_dtype)

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def code(self):
        return self.CODES[self.np_dtype]

    @property
    def size(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1

    @property
    def ndim(self):
        return len(self.shape)


def native_input(val):
    """Host array-like -> C-contiguous ndarray in the element type that crosses PCIe: uint8 / int8 / float16 /
    float32 stay as they are, everything else is cast to float32 on the host like the reference (`Parameter.py:13`)."""
    a = np.asarray(val)
    if a.dtype not in RawInput.CODES:
        a = a.astype(np.float32)
    return a


def is_device(x):
    return isinstance(x, DeviceArray)
