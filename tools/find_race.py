"""Developer tool: locate the first node whose output differs between replays of the same captured graph.

    python tools/find_race.py [--workload googlenet-v1] [--batch 256] [--iters 60]

The network is loaded with reuse_buffers=False, so every node output keeps its own arena address and can be digested
after each replay.  Prints, per replay that differs from the first, the first nodes (in schedule order) with a new digest.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyopenvino_b200.device import is_device  # noqa: E402
from pyopenvino_b200.inference_engine import IECore  # noqa: E402
from tools.synth_bin import ensure_model, synth_input  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='googlenet-v1', choices=sorted(bench.WORKLOADS))
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--iters', type=int, default=60)
ap.add_argument('--seed', type=int, default=31)
args = ap.parse_args()
model, _, _ = bench.WORKLOADS[args.workload]
xml = ensure_model(model, bench.CACHE)
ie = IECore()
net = ie.read_network(xml, xml[:-4] + '.bin')
exe = ie.load_network(net, 'B200', batch_size=args.batch, reuse_buffers=False)
in_name = net.inputs[0]['name']
x = synth_input(model, batch=args.batch, seed=args.seed)
for _ in range(3):
    exe.infer({in_name: x})
G = exe.ienet.G
tensors = []
for n in exe.task_list:
    node = G.nodes[n]
    for port, info in node.get('output', {}).items():
        d = info.get('data')
        if d is not None and is_device(d) and node['type'] != 'Const':
            tensors.append((node['name'], node['type'], d.t))
print('{} device node outputs'.format(len(tensors)))


def digest():
    torch.cuda.synchronize()
    return [int(t.view(torch.int32).to(torch.int64).sum().item()) for _, _, t in tensors]


ref = digest()
bad = 0
for it in range(args.iters):
    exe.infer({in_name: x})
    d = digest()
    diff = [i for i, (a, b) in enumerate(zip(ref, d)) if a != b]
    if diff:
        bad += 1
        print('replay {}: {} outputs differ; first: {}'.format(it, len(diff), [(tensors[i][0][-45:], tensors[i][1]) for i in diff[:4]]))
print('{} batch {}: {} of {} replays differ from the first'.format(args.workload, args.batch, bad, args.iters))
