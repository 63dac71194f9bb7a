"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: batch sharding, weight broadcast, output
gather, max-over-ranks timing reduction."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import REPO


def test_shard_range_covers_batch_without_overlap():
    from pyopenvino_b200.distributed import shard_range
    for total in (0, 1, 7, 8, 64, 257):
        for size in (1, 2, 3, 8):
            spans = [shard_range(total, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            lens = [e - b for b, e in spans]
            assert max(lens) - min(lens) <= 1


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, {repo!r})
    import numpy as np, torch
    from pyopenvino_b200 import distributed as D
    rank, size, local = D.init(backend='gloo')
    assert size == 2
    # weights: rank 0 is the source of truth
    flat = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
    D.broadcast_weights(flat)
    assert torch.equal(flat, torch.arange(1000, dtype=torch.float32))
    # batch sharding + gather, even split: global batch 6 x 10 "probabilities"
    full = torch.arange(60, dtype=torch.float32).view(6, 10)
    b, e = D.shard_range(6, rank, size)
    out = D.gather_outputs(full[b:e].clone())
    assert torch.equal(out, full)
    # uneven split
    full7 = torch.arange(70, dtype=torch.float32).view(7, 10)
    spans = [D.shard_range(7, r, size) for r in range(size)]
    b, e = spans[rank]
    out = D.gather_outputs(full7[b:e].clone(), sizes=[s[1] - s[0] for s in spans])
    assert torch.equal(out, full7)
    # step time reported = slowest rank
    assert D.max_over_ranks(1.0 + rank) == 2.0
    D.barrier()
    print('rank', rank, 'ok')
''')


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(repo=REPO))
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES='')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert 'rank {} ok'.format(rank) in out
