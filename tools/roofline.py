"""Algorithmic work per fused layer, from the IR dims.

FLOPs = 2*MACs for Convolution / GroupConvolution / MatMul; bytes (FP32) = 4*(numel(in)+numel(out))
per non-folded layer + 4*numel(W); folded elementwise nodes and in-place Concat count 0.
Measurement tooling only (used by bench.py and the profile summaries).
"""
import numpy as np


def _numel(dims):
    return int(np.prod(dims)) if len(dims) else 1


def layer_work(exe):
    """{node_id: {'flops', 'bytes', 'kind'}} for every live (non-folded, non-Const) step of exe's plan."""
    G = exe.ienet.G
    plan = exe._plan if exe._plan is not None else exe.build_plan()
    work = {}
    for n in exe.task_list:
        node = G.nodes[n]
        t = node['type']
        if plan[n]['skip'] or t in ('Const', 'Result'):
            continue
        tail = G.nodes[plan[n]['store_as']]
        out_dims = tail['output'][next(iter(tail['output']))]['dims'] if 'output' in tail else ()
        ins = node.get('input', {})
        flops, nbytes, kind = 0, 0, 'glue'
        if t == 'Convolution':
            xd, wd = ins[0]['dims'], ins[1]['dims']
            flops = 2 * _numel(out_dims) * wd[1] * wd[2] * wd[3]
            nbytes = 4 * (_numel(xd) + _numel(out_dims) + _numel(wd))
            kind = 'conv{}x{}'.format(wd[2], wd[3])
        elif t == 'GroupConvolution':
            xd, wd = ins[0]['dims'], ins[1]['dims']
            flops = 2 * _numel(out_dims) * wd[3] * wd[4]
            nbytes = 4 * (_numel(xd) + _numel(out_dims) + _numel(wd))
            kind = 'depthwise'
        elif t == 'MatMul':
            ad, bd = ins[0]['dims'], ins[1]['dims']
            k = ad[0] if node['data']['transpose_a'] == 'true' else ad[1]
            flops = 2 * _numel(out_dims) * k
            nbytes = 4 * (_numel(ad) + _numel(out_dims) + _numel(bd))
            kind = 'matmul'
        elif t in ('MaxPool', 'AvgPool', 'LRN', 'SoftMax', 'Sigmoid', 'ReLU', 'Clamp', 'Add', 'Multiply'):
            nbytes = 4 * (_numel(ins[0]['dims']) + _numel(out_dims))
            kind = t.lower()
        elif t == 'Parameter':
            nbytes = 4 * 2 * _numel(out_dims)          # NCHW staging buffer -> NHWC (+ folded mean/scale)
            kind = 'input_layout'
        elif t == 'Concat':
            inplace = all(plan[h]['out_slot'] is not None and plan[h]['out_slot'][0] == n
                          for h in exe.task_list if plan[h].get('out_slot') and plan[h]['out_slot'][0] == n)
            slots = plan[n].get('concat_slots')
            copied = 0
            if slots is not None:
                writers = {plan[h]['store_as'] for h in exe.task_list if plan[h].get('out_slot') and plan[h]['out_slot'][0] == n}
                for src, (off, c) in slots.items():
                    if src not in writers:
                        copied += _numel(out_dims) // out_dims[1] * c
            else:
                copied = _numel(out_dims)
            nbytes = 8 * copied
            kind = 'concat'
        elif t in ('Transpose',):
            nbytes = 0        # NHWC-resident: metadata only for the [0,2,3,1] permutation
            kind = 'transpose'
        work[n] = {'flops': int(flops), 'bytes': int(nbytes), 'kind': kind, 'type': t, 'name': node['name']}
    # sibling 1x1 convolutions run as one contraction: their work is accounted on the group head (the shared
    # input is read once)
    for n in list(work):
        head = plan[n].get('grouped')
        if head is not None and head in work:
            x_bytes = 4 * _numel(G.nodes[n]['input'][0]['dims'])
            work[head]['flops'] += work[n]['flops']
            work[head]['bytes'] += work[n]['bytes'] - x_bytes
            work[head]['name'] += ' + ' + work[n]['name'].split('/')[-2] if '/' in work[n]['name'] else ''
            del work[n]
    return work


def roofline_ms(w, hbm_gbs, tensor_tflops):
    return max(w['flops'] / (tensor_tflops * 1e12), w['bytes'] / (hbm_gbs * 1e9)) * 1e3
