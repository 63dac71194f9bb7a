"""Multi-GPU plumbing: one process per GPU, batch sharded across ranks, weights replicated.

The hot path has no cross-image dependency, so there is no data-path
collective: NCCL (over NVLink / NVSwitch, through `torch.distributed`) is used only to broadcast
the weight arena from rank 0 once per `load_network` and to gather the per-rank Result tensors.
On a CPU-only box the same code runs over gloo (used by the world_size-2 tests).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> (0, 1, 0))."""
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend=None):
    """Join the process group when launched under torchrun with WORLD_SIZE > 1."""
    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, size, local


def shard_range(total, rank, size):
    """[begin, end) of the images rank `rank` owns out of `total` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(total, size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_weights(flat, src=0):
    """Broadcast the weight arena (one flat tensor, see Executable_Network.load_constants) from `src`;
    afterwards every replica computes with bit-identical weights."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda:
            # the arena was filled on the engine's own stream and is read there again right after: order the
            # collective (queued relative to the CURRENT stream) against both sides with full device syncs
            torch.cuda.synchronize(flat.device)
        dist.broadcast(flat, src=src)
        if flat.is_cuda:
            torch.cuda.synchronize(flat.device)
    return flat


def gather_outputs(local, sizes=None):
    """All-gather per-rank result rows (dim 0 = images) into the global batch order.

    `local`: torch tensor [b_local, ...] (device tensor for NCCL, CPU tensor for gloo).  Ranks may own
    different numbers of rows when the batch does not divide evenly (`sizes` = rows per rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    size = dist.get_world_size()
    if sizes is None or len(set(sizes)) == 1:
        out = torch.empty((size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # uneven shards: pad every rank's rows to the largest shard, gather, drop the padding
    rows = max(sizes)
    padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((size * rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * rows:r * rows + n] for r, n in enumerate(sizes)], dim=0)


def max_over_ranks(value):
    """Max of a python float over ranks (device-timed step durations are reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
