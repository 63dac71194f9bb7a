"""ORACLE (test infrastructure only) -- CPU restatement of the reference IR executor.

Restates `pyopenvino/inference_engine.py` of yas-sim/pyopenvino: IR XML + `.bin` parsing
(`:105-199`), the ready-list schedule (`:218-242`), the per-node run loop that hands host ndarrays
from node to node (`:245-292`) and `infer` (`:295-321`), dispatching to the numpy restatements in
`oracle/ref_ops.py`.  Only tests, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs may import this; the product never does.

Differences that are deliberate and stated:
  * `kernel_type` is 'numpy' or 'special' only (the 'naive' python loops are never an oracle,
    SURVEY.md section 8c);
  * `faithful_const=True` re-materialises every Const from a python tuple on every inference exactly
    like `Const.py:13` + `inference_engine.py:197-199` (that is 45% of the reference's GoogLeNet
    time and is what the CPU baseline times); `False` keeps ndarrays (same values, faster tests);
  * the reference is batch-1 only; `infer_batch` defines the batched meaning as the stack of
    independent batch-1 runs.
"""
import os
import struct
import time
import xml.etree.ElementTree as et

import numpy as np

from . import ref_ops as R

_FMT = {'FP32': ('f', 4), 'F32': ('f', 4), 'FP16': ('e', 2), 'F16': ('e', 2), 'I64': ('q', 8),
        'I32': ('i', 4), 'I16': ('h', 2), 'I8': ('b', 1), 'U8': ('B', 1)}   # common_def.py:13-14


class RefNetwork:
    """IENetwork restatement (inference_engine.py:94-207)."""

    def __init__(self, xml_path, faithful_const=False):
        base = os.path.splitext(xml_path)[0]
        if not os.path.isfile(base + '.xml') or not os.path.isfile(base + '.bin'):
            raise Exception('model {} is not found'.format(xml_path))
        root = et.parse(base + '.xml').getroot()
        if root.tag != 'net':
            raise Exception('Not an OpenVINO IR file')
        with open(base + '.bin', 'rb') as f:
            blob = f.read()
        self.nodes = {}
        for layer in root.findall('./layers/layer'):
            nid = int(layer.attrib['id'])
            node = {k: v for k, v in layer.attrib.items() if k != 'id'}
            data = layer.find('data')
            if data is not None:
                node['data'] = dict(data.attrib)
                for key in ('shape', 'stride'):
                    if key in node['data']:
                        txt = node['data'][key]
                        node['data'][key] = R.ints(txt) if txt.strip() else ()
            for tag in ('input', 'output'):
                ports = layer.find(tag)
                if ports is not None:
                    node[tag] = {}
                    for port in ports.findall('port'):
                        node[tag][int(port.attrib['id'])] = {
                            'precision': port.attrib['precision'],
                            'dims': tuple(int(d.text) for d in port.findall('dim'))}
            self.nodes[nid] = node
        self.edges = [(int(e.attrib['from-layer']), int(e.attrib['from-port']),
                       int(e.attrib['to-layer']), int(e.attrib['to-port']))
                      for e in root.findall('./edges/edge')]
        self.pred = {nid: [] for nid in self.nodes}
        for fl, fp, tl, tp in self.edges:
            self.pred[tl].append((fl, fp, tp))
        # constants (inference_engine.py:188-199)
        for nid, node in self.nodes.items():
            if node['type'] != 'Const':
                continue
            d = node['data']
            off, size = int(d['offset']), int(d['size'])
            prec = d['element_type'].upper()
            raw = blob[off:off + size]
            if faithful_const:
                ch, width = _FMT[prec]
                node['const'] = struct.unpack('<' + ch * (len(raw) // width), raw)
            else:
                node['const'] = np.frombuffer(raw, dtype=R.DTYPES[d['element_type']]).copy()
        self.inputs = [n for n in self.nodes.values() if n['type'] == 'Parameter']
        self.outputs = [n for n in self.nodes.values() if n['type'] == 'Result']
        self.schedule()

    def schedule(self):
        """inference_engine.py:218-242: sources first, then repeated ready sweeps."""
        done, order, pending = set(), [], []
        for nid, node in self.nodes.items():
            if node['type'] in ('Const', 'Parameter'):
                order.append(nid)
                done.add(nid)
            else:
                pending.append(nid)
        while pending:
            rest = []
            for nid in pending:
                if all(p[0] in done for p in self.pred[nid]):
                    order.append(nid)
                    done.add(nid)
                else:
                    rest.append(nid)
            assert len(rest) < len(pending), 'graph is not a DAG'
            pending = rest
        self.order = order


def _first_out(node):
    return next(iter(node['output']))


def run_node(node, ins, kernel_type):
    """Dispatch one node to its ref_ops restatement; returns {out_port: ndarray} (or {} for Result)."""
    t = node['type']
    d = node.get('data', {})
    for port, arr in ins.items():        # validation convention, e.g. Convolution.py:154-157
        spec = node['input'][port]
        assert arr.dtype == R.DTYPES[spec['precision']], (node['name'], arr.dtype, spec)
        assert arr.shape == spec['dims'], (node['name'], arr.shape, spec['dims'])
    if t == 'Const':
        return {0: np.array(node['const'], dtype=R.DTYPES[d['element_type']]).reshape(d['shape'])}
    if t == 'Parameter':
        return {0: np.array(node['param']).reshape(d['shape']).astype(R.DTYPES[d['element_type']])}
    if t == 'Result':
        node['result'] = ins[0]
        return {}
    op = _first_out(node)
    if t == 'Convolution':
        res = R.convolution(d, ins[0], ins[1], kernel_type, R.DTYPES[node['output'][op]['precision']])
    elif t == 'GroupConvolution':
        res = R.group_convolution(d, ins[0], ins[1])
    elif t == 'MatMul':
        res = R.matmul(d, ins[0], ins[1])
    elif t == 'MaxPool':
        res = R.maxpool(d, ins[0])
    elif t == 'AvgPool':
        res = R.avgpool(d, ins[0])
    elif t == 'Add':
        res = R.add(ins[0], ins[1])
    elif t == 'Multiply':
        res = R.multiply(ins[0], ins[1])
    elif t == 'ReLU':
        res = R.relu(ins[0])
    elif t == 'Clamp':
        res = R.clamp(d, ins[0])
    elif t == 'SoftMax':
        res = R.softmax(ins[0])
    elif t == 'Sigmoid':
        res = R.sigmoid(ins[0])
    elif t == 'LRN':
        res = R.lrn(d, ins[0])
    elif t == 'Concat':
        res = R.concat(d, ins.values())
    elif t == 'Transpose':
        res = R.transpose(ins[0], ins[1])
    elif t == 'Reshape':
        res = R.reshape(ins[0], ins[1])
    elif t == 'Unsqueeze':
        res = R.unsqueeze(ins[0], ins[1])
    elif t == 'ShapeOf':
        res = R.shape_of(node['input'][0]['dims'], R.DTYPES[node['output'][op]['precision']])
    elif t == 'StridedSlice':
        res = R.strided_slice(ins[0], ins[1], ins[2], ins[3])
    elif t == 'PriorBoxClustered':
        res = R.prior_box_clustered(d, ins[0], ins[1])
    elif t == 'DetectionOutput':
        res = R.detection_output(d, ins[0], ins[1], ins[2])
    else:
        raise SystemExit("ERROR: Operation '{}' (node={}) is not supported.".format(t, node['name']))
    return {op: res}


class RefExecutable:
    """Executable_Network restatement (inference_engine.py:211-321)."""

    def __init__(self, net, kernel_type='numpy'):
        self.net = net
        self.kernel_type = kernel_type
        self.node_seconds = {}

    def infer(self, inputs, keep=False):
        net = self.net
        for name, val in inputs.items():
            for node in net.nodes.values():
                if node['name'] == name:
                    node['param'] = val
        for nid in net.order:
            node = net.nodes[nid]
            ins = {}
            for fl, fp, tp in net.pred[nid]:     # inference_engine.py:245-256
                ins[tp] = net.nodes[fl]['output'][fp]['data']
            t0 = time.perf_counter()
            res = run_node(node, ins, self.kernel_type)
            self.node_seconds[node['type']] = self.node_seconds.get(node['type'], 0.0) + time.perf_counter() - t0
            for port, arr in res.items():
                node['output'][port]['data'] = arr
        return {n['name']: n['result'] for n in net.outputs}

    def node_outputs(self):
        """{node_name: first output ndarray} of the last inference (the reference keeps every
        feature map alive in the graph, inference_engine.py:290-292)."""
        out = {}
        for node in self.net.nodes.values():
            if 'output' in node:
                port = _first_out(node)
                if 'data' in node['output'][port]:
                    out[node['name']] = node['output'][port]['data']
        return out

    def infer_batch(self, name, batch):
        """Batched meaning = stack of independent batch-1 results."""
        outs = {}
        for i in range(batch.shape[0]):
            r = self.infer({name: batch[i:i + 1]})
            for k, v in r.items():
                outs.setdefault(k, []).append(np.array(v))
        return {k: np.concatenate(v, axis=0) for k, v in outs.items()}


def load(xml_path, kernel_type='numpy', faithful_const=False):
    return RefExecutable(RefNetwork(xml_path, faithful_const), kernel_type)
