"""Replay determinism check: the same batch through the captured graph N times must give bit-identical results
(the contraction kernel's accumulation order is fixed; a protocol race between its warp roles would show up here).

    python tools/determinism.py [--workload googlenet-v1] [--batch 64] [--iters 200]
"""
import argparse
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyopenvino_b200.inference_engine import IECore  # noqa: E402
from tools.synth_bin import ensure_model, synth_input  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='googlenet-v1', choices=sorted(bench.WORKLOADS))
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--iters', type=int, default=200)
args = ap.parse_args()
model, _, _ = bench.WORKLOADS[args.workload]
xml = ensure_model(model, bench.CACHE)
ie = IECore()
net = ie.read_network(xml, xml[:-4] + '.bin')
exe = ie.load_network(net, 'B200', batch_size=args.batch)
in_name = net.inputs[0]['name']
x = synth_input(model, batch=args.batch, seed=7)
digests = set()
for i in range(args.iters):
    out = exe.infer({in_name: x})
    h = hashlib.sha256()
    for k in sorted(out):
        h.update(np.ascontiguousarray(out[k]).tobytes())
    digests.add(h.hexdigest())
print('{} batch {}: {} inferences, {} distinct result digest(s)'.format(args.workload, args.batch, args.iters, len(digests)))
sys.exit(0 if len(digests) == 1 else 1)
