"""Developer tool: first replay vs second replay of a FRESHLY loaded network, repeated, with the arena carved out of memory
that was filled with a poison pattern just before (NaN by default): anything that reads bytes this inference did not write
shows up as a difference between the two replays (or as NaN / a range fallback).

    python tools/find_race_fresh.py [--workload googlenet-v1] [--batch 256] [--trials 12] [--poison nan|big|none] [--reuse 0|1]
"""
import argparse
import gc
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyopenvino_b200.device import is_device  # noqa: E402
from pyopenvino_b200.inference_engine import IECore  # noqa: E402
from tools.synth_bin import ensure_model, synth_input  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='googlenet-v1', choices=sorted(bench.WORKLOADS))
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--trials', type=int, default=12)
ap.add_argument('--poison', default='nan')
ap.add_argument('--reuse', type=int, default=0)
ap.add_argument('--gb', type=float, default=20.0)
args = ap.parse_args()
model, _, _ = bench.WORKLOADS[args.workload]
xml = ensure_model(model, bench.CACHE)
x = synth_input(model, batch=args.batch, seed=31)
bad = 0
for trial in range(args.trials):
    if args.poison != 'none':
        junk = torch.empty(int(args.gb * (1 << 30)) // 4, dtype=torch.float32, device='cuda')
        junk.fill_(float('nan') if args.poison == 'nan' else 3.0e38)
        torch.cuda.synchronize()
        del junk                                   # back to torch's caching allocator: the next arena is carved out of it
    ie = IECore()
    net = ie.read_network(xml, xml[:-4] + '.bin')
    exe = ie.load_network(net, 'B200', batch_size=args.batch, reuse_buffers=bool(args.reuse))
    in_name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
    r1 = exe.infer({in_name: x})[out_name]
    G = exe.ienet.G
    tensors = []
    if not args.reuse:
        for n in exe.task_list:
            node = G.nodes[n]
            for port, info in node.get('output', {}).items():
                d = info.get('data')
                if d is not None and is_device(d) and node['type'] != 'Const':
                    tensors.append((node['name'], node['type'], d.t))
    torch.cuda.synchronize()
    d1 = [t.view(torch.int32).to(torch.int64).sum().item() for _, _, t in tensors]
    r2 = exe.infer({in_name: x})[out_name]
    torch.cuda.synchronize()
    d2 = [t.view(torch.int32).to(torch.int64).sum().item() for _, _, t in tensors]
    r3 = exe.infer({in_name: x})[out_name]
    same12, same23 = np.array_equal(r1, r2), np.array_equal(r2, r3)
    diff = [i for i, (a, b) in enumerate(zip(d1, d2)) if a != b]
    msg = ''
    if not same12 or not same23 or diff:
        bad += 1
        rows = np.unique(np.nonzero(r1 != r2)[0])
        msg = ' rows differing (1 vs 2): {} max|d| {:.3g}; first nodes: {}'.format(
            rows[:8].tolist(), float(np.nanmax(np.abs(r1 - r2))) if rows.size else 0.0,
            [(tensors[i][0][-40:], tensors[i][1]) for i in diff[:4]])
    print('trial {}: replay1==replay2 {} replay2==replay3 {} nan {} fallbacks {}{}'.format(
        trial, same12, same23, bool(np.isnan(r1).any()), getattr(exe, 'range_fallbacks', 0), msg), flush=True)
    del exe, net, ie, tensors
    gc.collect()
    torch.cuda.empty_cache() if args.poison == 'none' else None
print('{} of {} trials differ'.format(bad, args.trials))
