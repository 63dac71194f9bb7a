#!/usr/bin/env python
"""Concurrent host->device copy ceiling of the box: N ranks doing NOTHING but pinned cudaMemcpyAsync H2D.

    python tools/h2d_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/h2d_ceiling.py [--mb 154.1] [--iters 40]

Every rank copies a `--mb` MB pinned buffer (default: one GoogLeNet batch-256 FP32 input, 154.1 MB; also the uint8
form, 38.5 MB) to its GPU `--iters` times back to back through libb200ov's b200ov_memcpy_h2d (cudaMemcpyAsync on a
stream), all ranks between two barriers.  Rank 0 prints one JSON line: per-rank GB/s (min / median / max) and the
aggregate.  This is the denominator of bench.py's `e2e` at N GPUs: an engine cannot ingest FP32 images faster than
the host can feed them.  D2H and bidirectional copies are measured the same way.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mb', type=float, nargs='*', default=[154.14, 38.54, 8.0])
    ap.add_argument('--iters', type=int, default=40)
    args = ap.parse_args()
    import torch
    from pyopenvino_b200 import _cabi, device, distributed
    rank, world, local = distributed.init()
    torch.cuda.set_device(local)
    device.init(local)
    if world > 1:
        device.bind_host_thread(local, world)
    stream = torch.cuda.Stream()
    s = C.c_void_p(stream.cuda_stream)
    rows = []
    for mb in args.mb:
        n = int(mb * 1e6)
        host = device.pinned_empty(n, torch.uint8, zero=True)
        host2 = device.pinned_empty(n, torch.uint8, zero=True)
        dev_in = torch.empty(n, dtype=torch.uint8, device='cuda')
        dev_out = torch.ones(n, dtype=torch.uint8, device='cuda')
        stream2 = torch.cuda.Stream()
        s2 = C.c_void_p(stream2.cuda_stream)
        for direction in ('h2d', 'd2h', 'both'):
            def once():
                if direction in ('h2d', 'both'):
                    _cabi.call('b200ov_memcpy_h2d', C.c_void_p(dev_in.data_ptr()), C.c_void_p(host.data_ptr()), C.c_size_t(n), s)
                if direction in ('d2h', 'both'):
                    _cabi.call('b200ov_memcpy_d2h', C.c_void_p(host2.data_ptr()), C.c_void_p(dev_out.data_ptr()), C.c_size_t(n), s2)
            for _ in range(3):
                once()
            torch.cuda.synchronize()
            distributed.barrier()
            t0 = time.perf_counter()
            for _ in range(args.iters):
                once()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            distributed.barrier()
            gbs = n * args.iters / dt / 1e9
            if world > 1:
                t = torch.tensor([gbs], dtype=torch.float64, device='cuda')
                allg = [torch.zeros_like(t) for _ in range(world)]
                torch.distributed.all_gather(allg, t)
                per = sorted(float(v.item()) for v in allg)
            else:
                per = [gbs]
            rows.append({'mb': mb, 'dir': direction, 'per_rank_gbs_min': per[0], 'per_rank_gbs_median': per[len(per) // 2],
                         'per_rank_gbs_max': per[-1], 'aggregate_gbs': sum(per), 'note': 'per direction' if direction == 'both' else ''})
    if rank == 0:
        print(json.dumps({'tool': 'h2d_ceiling', 'n_gpus': world, 'iters': args.iters, 'host_cpus': os.cpu_count(),
                          'numa_bound_cpus': sorted(device._numa_cpus) if device._numa_cpus else None, 'rows': rows}), flush=True)
    distributed.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
