"""pyopenvino_b200 inference engine: the reference's `IECore / read_network / load_network / infer`
surface (`pyopenvino/inference_engine.py`) over HBM-resident feature maps and libb200ov kernels.

Kept from the reference:
  * `IECore()` discovers one plugin module per IR layer type, file basename == type string
    (`inference_engine.py:23-43`); every node is executed through
    `plugins[type].compute(node, inputs, kernel_type=..., debug=False)` (`:280`).
  * `read_network(model, weights)` -> `IENetwork` with `.G` (networkx.DiGraph, node-attribute schema
    of `README.md:88-125`, edges carry `connection=(from, fport, to, tport)`), `.inputs`, `.outputs`,
    `.layers`, `.edges` (`:94-207`).
  * `load_network(network, device_name, num_requests)` -> `Executable_Network` with
    `.kernel_type`, `.expected_result`, `.pickle_node_args`, `.schedule_tasks()`,
    `.infer({name: array}, verbose) -> {result_name: ndarray}` (`:211-321`).

New, B200-side behaviour:
  * what flows along the edges is a `DeviceArray` (NHWC in HBM), not a host ndarray; the only
    host<->device copies are at Parameter and Result.
  * `load_network(..., batch_size=B)` re-batches the static IR (the reference is batch-1 only); the
    batched meaning is "B independent batch-1 results" (SURVEY.md section 0.4).
  * a fusion plan folds Add / ReLU / Clamp / Multiply nodes into producer epilogues and lets Concat
    producers write in place; the whole launch sequence is captured once into a CUDA graph and
    replayed (`fuse`, `use_graph`).  `fuse=False, use_graph=False` gives the reference's node-by-node
    behaviour with every node output materialised (used for per-node parity tests).
There is no CPU execution path: `device_name` is accepted and ignored like in the reference
(`inference_engine.py:86-90`), and everything runs on the current CUDA device.
"""
import glob
import importlib
import os
import pickle
import sys
import time
import xml.etree.ElementTree as et

import networkx as nx
import numpy as np

from . import common_def

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_PLUGIN_DIR = os.path.join(_PKG_DIR, 'op_plugins')
_PLUGIN_PKG = __name__.rsplit('.', 1)[0] + '.op_plugins'


# -------------------------------------------------------------------------------------------------

class Plugins:
    """Operator-plugin registry: module file basename == IR layer type."""

    def __init__(self):
        self.plugins = {}

    def import_plugin(self, plugin_path: str, file_path: str, plugin_name: str = None):
        bname = os.path.splitext(os.path.basename(file_path))[0]
        if plugin_name is None:
            plugin_name = bname
        module = importlib.import_module(plugin_path.replace('/', '.') + '.' + bname)
        setattr(self, plugin_name, module)
        self.plugins[plugin_name] = module

    def load_plugins(self, plugin_path: str = None):
        """Import every `<Type>.py` under the plugin directory (default: this package's op_plugins)."""
        directory = _PLUGIN_DIR if plugin_path is None else plugin_path
        package = _PLUGIN_PKG if plugin_path is None else plugin_path
        for path in sorted(glob.glob(os.path.join(directory, '**', '*.py'), recursive=True)):
            if os.path.basename(path).startswith('_'):
                continue
            self.import_plugin(package, path)


class IECore:
    def __init__(self):
        self.plugins = Plugins()
        self.plugins.load_plugins()
        common_def.enable_escape_sequence()

    def construct_node_info(self, net, node_type: str) -> list:
        return [net.G.nodes[node_id] for node_id, _ in net.find_node_by_type(node_type)]

    def check_nodes(self, G: nx.DiGraph):
        unsupported = {G.nodes[n]['type'] for n in G.nodes if G.nodes[n]['type'] not in self.plugins.plugins}
        if unsupported:
            print('\x1b[31mUnsupported nodes : {}\x1b[37m'.format(unsupported))
        return unsupported

    # OpenVINO Inference Engine API
    def read_network(self, model: str, weights: str = None):
        net = IENetwork(self)
        net.read_IR_Model(model)
        net.parse_IR_XML()
        net.build_graph()
        net.set_constants_to_graph()
        net.inputs = self.construct_node_info(net, 'Parameter')
        net.outputs = self.construct_node_info(net, 'Result')
        return net

    # OpenVINO Inference Engine API
    def load_network(self, network, device_name: str = 'B200', num_requests: int = 1, batch_size: int = None,
                     fuse: bool = True, use_graph: bool = True, reuse_buffers: bool = True, storage: str = 'f32'):
        """`storage='f16'` (opt-in, SURVEY.md 8(f) NEXT-4): NHWC feature maps between nodes are kept in HBM as FP16 -- half
        the bytes of every bandwidth-bound layer, and a contraction whose input already is FP16 needs 2 instead of 3
        tensor-core MMAs per product.  Arithmetic stays FP32 (FP32 accumulate, round to nearest even on store); network
        inputs, weights, biases, 2-D tensors and results stay FP32.  Results then carry FP16's 2^-11 relative rounding per
        stored feature map (tolerances: tests/test_gpu_f16_storage.py); the default 'f32' is the reference's precision."""
        assert storage in ('f32', 'f16')
        if batch_size is not None and batch_size != network.batch_size:
            network.set_batch_size(batch_size)
        exenet = Executable_Network(network, fuse=fuse, use_graph=use_graph, reuse_buffers=reuse_buffers, storage=storage)
        self.check_nodes(exenet.ienet.G)
        exenet.schedule_tasks()
        return exenet


# -------------------------------------------------------------------------------------------------

class IENetwork:
    def __init__(self, iecore: IECore):
        self.ie = iecore
        self.xml = None
        self.bin = None
        self.G = None
        self.layers = None
        self.edges = None
        self.inputs = None
        self.outputs = None
        self.batch_size = 1

    def read_IR_Model(self, model):
        bname, ext = os.path.splitext(model)
        xml_file, bin_file = bname + '.xml', bname + '.bin'
        if not os.path.isfile(xml_file) or not os.path.isfile(bin_file):
            raise Exception('model {} is not found'.format(model))
        self.xml = et.parse(xml_file)
        with open(bin_file, 'rb') as f:
            self.bin = f.read()

    def parse_IR_XML(self):
        root = self.xml.getroot()
        if root.tag != 'net':
            raise Exception('Not an OpenVINO IR file')
        layers = {}
        for layer in root.findall('./layers/layer'):
            info = {key: val for key, val in layer.attrib.items() if key != 'id'}
            data = layer.find('data')
            if data is not None:
                info['data'] = dict(data.attrib)
                for key in ('shape', 'stride'):
                    if key in info['data']:
                        text = info['data'][key]
                        info['data'][key] = common_def.string_to_tuple(text) if text.strip() else ()
            for tag in ('input', 'output'):
                ports = layer.find(tag)
                if ports is not None:
                    info[tag] = {}
                    for port in ports.findall('port'):
                        dims = tuple(int(dim.text) for dim in port.findall('./dim'))
                        info[tag][int(port.attrib['id'])] = {'precision': port.attrib['precision'], 'dims': dims}
            layers[int(layer.attrib['id'])] = info
        self.layers = layers
        self.edges = [(int(e.attrib['from-layer']), int(e.attrib['from-port']), int(e.attrib['to-layer']),
                       int(e.attrib['to-port'])) for e in root.findall('./edges/edge')]

    def build_graph(self):
        self.G = nx.DiGraph()
        for node_id, node_info in self.layers.items():
            self.G.add_node(node_id)
            for key, val in node_info.items():
                self.G.nodes[node_id][key] = val
        for edge in self.edges:
            self.G.add_edge(edge[0], edge[2])
            self.G.edges[(edge[0], edge[2])]['connection'] = edge
        assert nx.is_directed_acyclic_graph(self.G)

    def set_constants_to_graph(self):
        """Slice every Const blob out of the `.bin` as an ndarray view (the reference unpacks python
        tuples here, `inference_engine.py:188-199`)."""
        for node_id, _ in self.find_node_by_type('Const'):
            node = self.G.nodes[node_id]
            data = node['data']
            offset, size = int(data['offset']), int(data['size'])
            precision = data['element_type'].upper()
            dtype = np.dtype(common_def.type_convert_tbl[data['element_type']])
            decoded = np.frombuffer(self.bin, dtype=dtype, count=size // dtype.itemsize, offset=offset)
            node['const'] = {'data': decoded, 'element_info': precision, 'size': size,
                             'decode_info': common_def.format_config[precision]}

    def find_node_by_type(self, type: str) -> list:
        return [(n, self.G.nodes[n]['name']) for n in self.G.nodes() if self.G.nodes[n]['type'] == type]

    # ---- batching (new: the reference has no reshape API) ---------------------------------------
    def set_batch_size(self, batch: int):
        """Rewrite the static shapes for `batch` images per inference.

        Every tensor downstream of a Parameter gets dim 0 multiplied by batch / old batch.  ShapeOf
        outputs are static shape vectors and stop the propagation.  DetectionOutput emits
        (1, 1, N*keep_top_k, 7) (`DetectionOutput.py:232-237`), so its dim 2 scales instead.
        """
        assert batch >= 1 and self.batch_size >= 1
        G = self.G
        dynamic = set()
        stack = [n for n in G.nodes if G.nodes[n]['type'] == 'Parameter']
        while stack:
            n = stack.pop()
            if n in dynamic or G.nodes[n]['type'] == 'ShapeOf':
                continue
            dynamic.add(n)
            stack.extend(G.successors(n))
        old = self.batch_size

        def scale(dims, axis):
            dims = list(dims)
            assert dims[axis] % old == 0
            dims[axis] = dims[axis] // old * batch
            return tuple(dims)

        for n in dynamic:
            node = G.nodes[n]
            axis = 2 if node['type'] == 'DetectionOutput' else 0
            for port in node.get('output', {}).values():
                if len(port['dims']) > axis:
                    port['dims'] = scale(port['dims'], axis)
            if node['type'] == 'Parameter':
                node['data']['shape'] = scale(node['data']['shape'], 0)
        for fl, fp, tl, tp in self.edges:
            if fl in dynamic:
                G.nodes[tl]['input'][tp]['dims'] = G.nodes[fl]['output'][fp]['dims']
        self.batch_size = batch


# -------------------------------------------------------------------------------------------------

class _Liveness:
    """Liveness-planned reuse of per-inference buffers (SURVEY.md 8(f) NEXT-3; ownership rule `inference_engine.py:290-292`:
    the reference keeps every feature map until the next inference only because nothing frees them -- no caller reads them
    after `infer`).  Active in planned CUDA-graph mode only.

    Every arena chunk counts the consumer steps still to be queued over all graph nodes whose output lives in it (views --
    Reshape, NHWC Transpose, channel slices of an in-place Concat -- share their producer's chunk).  When the count reaches
    zero the chunk goes back to the arena and the next producer may take it.  The sequence of allocations and releases is a
    pure function of the plan, so the warm-up pass, the captured pass and every later capture see identical addresses."""

    def __init__(self, exe, arena):
        self.exe, self.arena = exe, arena
        self.pending = {}          # chunk base pointer -> consumer steps not yet queued
        self.node_chunk = {}       # graph node id -> chunk base pointer of its stored output
        self.pinned = set()        # chunks that must survive the pass (Result tensors)

    def _consumers(self, node_id):
        G, plan = self.exe.ienet.G, self.exe._plan
        # a MaxPool folded into its consumer ('pool_into') still reads its input -- through that consumer's step
        return sum(1 for s in G.successors(node_id)
                   if (not plan[s]['skip'] or plan[s].get('pool_into') is not None) and not plan[s].get('const_done'))

    def stored(self, node_id, data):
        """`data` was stored as the output of graph node `node_id`."""
        t = getattr(data, 't', None)
        if t is None or not self.arena.owns(t):
            return
        key = t.data_ptr()
        self.node_chunk[node_id] = key
        self.pending[key] = self.pending.get(key, 0) + self._consumers(node_id)

    def pin(self, data):
        t = getattr(data, 't', None)
        if t is not None and self.arena.owns(t):
            self.pinned.add(t.data_ptr())

    def consumed(self, task):
        """Step `task` has been queued: its inputs have one consumer less."""
        G = self.exe.ienet.G
        preds = list(G.pred[task])
        pooled = self.exe._plan[task]['ops'].get('pre_pool')
        if pooled is not None:
            preds += list(G.pred[pooled])
        for pred in preds:
            key = self.node_chunk.get(pred)
            if key is None:
                continue
            self.pending[key] -= 1
            if self.pending[key] <= 0 and key not in self.pinned:
                self._release(key)

    def sweep(self):
        """Chunks whose producers have no consumer at all (dead outputs) are released at once."""
        for key, cnt in list(self.pending.items()):
            if cnt <= 0 and key not in self.pinned:
                self._release(key)

    def _release(self, key):
        chunk = self.arena.live.get(key)
        if chunk is not None:
            self.arena.release(self.arena.blocks[chunk[0]][chunk[1]:chunk[1] + chunk[2]])
        self.pending.pop(key, None)


_EPILOGUE_HEADS = ('Convolution', 'GroupConvolution', 'MatMul')
_OUT_CAPABLE = ('Convolution', 'GroupConvolution', 'MaxPool', 'AvgPool', 'LRN', 'ReLU', 'Clamp', 'Sigmoid')


class Executable_Network:
    def __init__(self, ienetwork: IENetwork, fuse: bool = True, use_graph: bool = True, reuse_buffers: bool = True,
                 storage: str = 'f32'):
        self.ienet = ienetwork
        self.storage = storage
        self.reuse_buffers = reuse_buffers and os.environ.get('B200OV_NO_REUSE') != '1'
        self.expected_result = None     # {node_name: [precision, dims, ndarray]} feature-map ground truth (debug)
        self.kernel_type = 'naive'      # accepted for compatibility: every value runs the CUDA kernels
        self.pickle_node_args = []      # node ids whose (node, inputs) are pickled for node unit tests (eager mode)
        self.fuse = fuse
        self.use_graph = use_graph
        self.task_list = []
        self.stream = None
        self._plan = None
        self._graph = None
        self._graph_key = None
        self._captured = {}             # input-dtype signature -> {'graph', 'static_out', 'launches'}
        self._arena = None
        self._static_in = {}
        self._static_out = {}
        self._user_inputs = {}
        self._graph_launches = 0
        self._step_events = None
        self._weights = None
        self.last_node_seconds = {}

    # ---- scheduling (inference_engine.py:218-242) -------------------------------------------------
    def schedule_tasks(self):
        G = self.ienet.G
        done, pending = set(), []
        self.task_list = []
        for node_id in G.nodes:
            if G.nodes[node_id]['type'] in ('Const', 'Parameter'):
                self.task_list.append(node_id)
                done.add(node_id)
            else:
                pending.append(node_id)
        while pending:
            rest = []
            for node_id in pending:
                if all(p in done for p in G.predecessors(node_id)):
                    self.task_list.append(node_id)
                    done.add(node_id)
                else:
                    rest.append(node_id)
            assert len(rest) < len(pending)
            pending = rest
        self._plan = None
        self._graph = None
        self._graph_key = None
        self._captured = {}

    def prepare_inputs_for_task(self, task) -> dict:
        G = self.ienet.G
        inputs = {}
        for predecessor in G.pred[task]:
            fl, fp, tl, tp = G.edges[(predecessor, task)]['connection']
            inputs[tp] = G.nodes[fl]['output'][fp].get('data')
        return inputs

    # ---- fusion plan -------------------------------------------------------------------------------
    def _const_operand(self, consumer, producer):
        """(node_id of the Const feeding `consumer` on its other port, that Const's dims) or None."""
        G = self.ienet.G
        others = [p for p in G.pred[consumer] if p != producer]
        if len(others) != 1 or G.nodes[others[0]]['type'] != 'Const':
            return None
        return others[0]

    def _single_consumer(self, node_id):
        G = self.ienet.G
        # ShapeOf only reads the static port dims (ShapeOf.py:21), so it does not count as a data consumer
        succ = [n for n in G.successors(node_id) if G.nodes[n]['type'] != 'ShapeOf']
        return succ[0] if len(succ) == 1 else None

    def _per_channel(self, const_id, channels, like_dims):
        """True if the Const broadcasts per channel (or is a scalar) against a tensor of `like_dims`."""
        dims = tuple(self.ienet.G.nodes[const_id]['data']['shape'])
        size = int(np.prod(dims)) if len(dims) else 1
        if size == 1:
            return True
        if size != channels:
            return False
        if len(like_dims) == 4:
            return len(dims) == 4 and dims[1] == channels
        return dims[-1] == channels

    def build_plan(self):
        """Decide, per node, what is folded into whom.  Result: self._plan = {node_id: step} with
        step = {'skip': bool, 'fused': {...const node ids...}, 'store_as': node_id, 'concat': (cid, off)}."""
        G = self.ienet.G
        plan = {n: {'skip': False, 'ops': {}, 'store_as': n, 'out_slot': None} for n in self.task_list}
        if not self.fuse:
            self._plan = plan
            return plan
        absorbed = set()

        def absorb(head, node):
            plan[node]['skip'] = True
            absorbed.add(node)
            plan[head]['store_as'] = node

        for n in self.task_list:
            if n in absorbed:
                continue
            node = G.nodes[n]
            t = node['type']
            if 'output' not in node:
                continue
            out_dims = node['output'][common_def.first_output_port(node)]['dims']
            ops = plan[n]['ops']
            tail = n
            if t in _EPILOGUE_HEADS:
                channels = out_dims[1] if len(out_dims) == 4 else out_dims[-1]
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] == 'Add':
                    c = self._const_operand(nxt, tail)
                    conn = G.edges[(tail, nxt)]['connection']
                    if c is not None and conn[3] == 0 and self._per_channel(c, channels, out_dims) and \
                            int(np.prod(G.nodes[c]['data']['shape'])) == channels:
                        ops['bias'] = c
                        absorb(n, nxt)
                        tail = nxt
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] in ('ReLU', 'Clamp'):
                    nn = G.nodes[nxt]
                    ops['act'] = ('relu',) if nn['type'] == 'ReLU' else ('clamp', float(nn['data']['min']), float(nn['data']['max']))
                    absorb(n, nxt)
                    tail = nxt
            elif t == 'MaxPool':
                channels = out_dims[1]
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] == 'Multiply':
                    c = self._const_operand(nxt, tail)
                    if c is not None and int(np.prod(G.nodes[c]['data']['shape'])) == channels and \
                            self._per_channel(c, channels, out_dims):
                        ops['scale'] = c
                        absorb(n, nxt)
                        tail = nxt
                        nxt2 = self._single_consumer(tail)
                        if nxt2 is not None and G.nodes[nxt2]['type'] == 'Add':
                            c2 = self._const_operand(nxt2, tail)
                            conn = G.edges[(tail, nxt2)]['connection']
                            if c2 is not None and conn[3] == 0 and int(np.prod(G.nodes[c2]['data']['shape'])) == channels \
                                    and self._per_channel(c2, channels, out_dims):
                                ops['shift'] = c2
                                absorb(n, nxt2)
                                tail = nxt2
            elif t == 'Parameter' and len(out_dims) == 4:
                channels = out_dims[1]
                ops['to_nhwc'] = True
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] == 'Multiply':
                    c = self._const_operand(nxt, tail)
                    if c is not None and self._per_channel(c, channels, out_dims):
                        ops['scale'] = c
                        absorb(n, nxt)
                        tail = nxt
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] == 'Add':
                    c = self._const_operand(nxt, tail)
                    conn = G.edges[(tail, nxt)]['connection']
                    if c is not None and conn[3] == 0 and self._per_channel(c, channels, out_dims):
                        ops['shift'] = c
                        absorb(n, nxt)
                        tail = nxt
                # The stem: when the only consumer is a Convolution that runs on the 8-channel super-pixel view of a
                # <= 4-channel image (even width, even horizontal stride), the layout kernel writes the contraction's
                # FP16 (hi, lo) pairs directly -- the split then costs once per pixel instead of once per filter tap in
                # the stem's producer warps (its limiter: 650 cycles per slot against 343 of MMA time, round 1).
                nxt = self._single_consumer(tail)
                if nxt is not None and G.nodes[nxt]['type'] == 'Convolution' and channels <= 4 and os.environ.get('B200OV_NO_SPLIT_INPUT') != '1':
                    cd = G.nodes[nxt]['data']
                    if G.edges[(tail, nxt)]['connection'][3] == 0 and common_def.string_to_tuple(cd['strides'])[1] % 2 == 0 and \
                            out_dims[3] % 2 == 0 and common_def.string_to_tuple(cd['dilations']) == (1, 1):
                        ops['split'] = True
        # MaxPool 3x3 / stride 1 / pads 1 whose only consumer is a 1x1 convolution (the pool -> pool_proj pair of an inception
        # module): the convolution's A producers take the 9-tap max on the way in (b200ov_conv_desc.pre_pool), so the pooled
        # tensor is neither written nor re-read.  FP32 storage only; bit-identical to the two separate kernels.
        if self.storage == 'f32' and os.environ.get('B200OV_NO_POOL_FUSE') != '1':
            for n in self.task_list:
                node = G.nodes[n]
                if node['type'] != 'MaxPool' or plan[n]['skip'] or plan[n]['ops']:
                    continue
                d = node['data']
                in_dims = node['input'][0]['dims']
                out_dims = node['output'][common_def.first_output_port(node)]['dims']
                if len(in_dims) != 4 or tuple(in_dims) != tuple(out_dims) or in_dims[1] % 8 != 0 or in_dims[3] > 63 or \
                        common_def.string_to_tuple(d['kernel']) != (3, 3) or common_def.string_to_tuple(d['strides']) != (1, 1) or \
                        common_def.string_to_tuple(d['pads_begin']) != (1, 1) or common_def.string_to_tuple(d['pads_end']) != (1, 1):
                    continue
                nxt = self._single_consumer(n)
                if nxt is None or G.out_degree(n) != 1 or G.nodes[nxt]['type'] != 'Convolution' or plan[nxt]['skip']:
                    continue
                cd = G.nodes[nxt]['data']
                wdims = G.nodes[nxt]['input'][1]['dims']
                if G.edges[(n, nxt)]['connection'][3] != 0 or tuple(wdims[2:]) != (1, 1) or \
                        common_def.string_to_tuple(cd['strides']) != (1, 1) or common_def.string_to_tuple(cd['pads_begin']) != (0, 0) or \
                        common_def.string_to_tuple(cd['pads_end']) != (0, 0) or common_def.string_to_tuple(cd['dilations']) != (1, 1):
                    continue
                plan[n]['skip'] = True
                plan[n]['pool_into'] = nxt
                plan[nxt]['ops']['pre_pool'] = n
        # Concat in place: producers whose only consumer is a channel Concat write into its buffer
        for n in self.task_list:
            node = G.nodes[n]
            if node['type'] != 'Concat' or int(node['data']['axis']) != 1:
                continue
            out_dims = node['output'][common_def.first_output_port(node)]['dims']
            if len(out_dims) != 4:
                continue
            # The Concat plugin sees its inputs in G.pred order (= edge insertion order, like the reference's
            # `inputs.values()`, Concat.py:12); the in-place layout is only right when that IS the IR port order and
            # every port has its own producer (a producer feeding two ports collapses into one DiGraph edge).
            conns = [G.edges[(pred, n)]['connection'] for pred in G.pred[n]]
            ports = [tp for (_fl, _fp, _tl, tp) in conns]
            if ports != sorted(ports) or len(conns) != len(node['input']):
                continue                 # fall back to the copying Concat
            slots = {}
            off = 0
            for fl, fp, tl, tp in conns:
                c = G.nodes[fl]['output'][fp]['dims'][1]
                slots[fl] = (off, c)
                off += c
            plan[n]['concat_slots'] = slots
            for tail_id, (coff, c) in slots.items():
                head = next((h for h in self.task_list if plan[h]['store_as'] == tail_id and not plan[h]['skip']), None)
                if head is None or G.out_degree(tail_id) != 1 or G.nodes[head]['type'] not in _OUT_CAPABLE:
                    continue
                if G.nodes[head]['type'] in ('ReLU', 'Clamp', 'Sigmoid', 'LRN') and head != tail_id:
                    continue
                plan[head]['out_slot'] = (n, coff, c)
        # Contraction -> contraction edges (3x3_reduce -> 3x3, 5x5_reduce -> 5x5, the 1x1 -> 3x3/s2 pairs of the SSD extra
        # layers): the producer's epilogue leaves the tensor in the consumer's operand form (FP16 hi / scaled-lo pairs, the same
        # 4 bytes per value), so the consumer's A producers only route words instead of splitting every value once per filter
        # tap and column tile.  Same bits downstream; FP32 storage only; never for a tensor anything else reads.
        if self.storage == 'f32' and os.environ.get('B200OV_NO_HL') != '1':
            for n in self.task_list:
                node = G.nodes[n]
                if node['type'] not in ('Convolution', 'GroupConvolution') or plan[n]['skip'] or plan[n]['out_slot'] is not None or \
                        'output' not in node:
                    continue
                tail = plan[n]['store_as']
                cout = node['output'][common_def.first_output_port(node)]['dims'][1]
                readers = list(G.successors(tail))
                if cout % 8 != 0 or not readers:
                    continue
                ok = True
                for r in readers:
                    rn = G.nodes[r]
                    if rn['type'] != 'Convolution' or plan[r]['skip'] or G.edges[(tail, r)]['connection'][3] != 0 or \
                            'pre_pool' in plan[r]['ops'] or common_def.string_to_tuple(rn['data']['dilations']) != (1, 1):
                        ok = False
                        break
                if ok:
                    plan[n]['ops']['hl_out'] = True
        # Sibling 1x1 convolutions (same input, stride 1, no padding, bias + the same activation): one contraction
        # with several output tensors -- the 1x1 / 3x3_reduce / 5x5_reduce branches of an inception module read
        # (and FP16-split) their common input once instead of three times.
        groups = {}
        for n in (self.task_list if os.environ.get('B200OV_NO_GROUP') != '1' else []):
            node = G.nodes[n]
            if node['type'] != 'Convolution' or plan[n]['skip']:
                continue
            d = node['data']
            wdims = node['input'][1]['dims']
            if tuple(wdims[2:]) != (1, 1) or common_def.string_to_tuple(d['strides']) != (1, 1) or \
                    common_def.string_to_tuple(d['pads_begin']) != (0, 0) or common_def.string_to_tuple(d['pads_end']) != (0, 0) or \
                    common_def.string_to_tuple(d['dilations']) != (1, 1) or wdims[1] % 8 != 0:
                continue
            src = next(((fl, fp) for fl in G.pred[n] for (_fl, fp, _tl, tp) in [G.edges[(fl, n)]['connection']] if tp == 0), None)
            wsrc = next((fl for fl in G.pred[n] if G.edges[(fl, n)]['connection'][3] == 1), None)
            if src is None or wsrc is None or G.nodes[wsrc]['type'] != 'Const':
                continue
            groups.setdefault((src, plan[n]['ops'].get('act'), 'bias' in plan[n]['ops']), []).append(n)
        for members in groups.values():
            for i in range(0, len(members) - 1, 3):
                chunk = members[i:i + 3]
                if len(chunk) >= 2:
                    plan[chunk[0]]['group'] = chunk
                    for m in chunk[1:]:
                        plan[m]['grouped'] = chunk[0]
        self._plan = plan
        return plan

    # ---- execution ---------------------------------------------------------------------------------
    def _ensure_device(self):
        from . import device as dev
        import torch
        dev.init()
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        return dev

    def _run(self, verbose=False, capture=False):
        """One pass over the task list.  In planned mode folded nodes are skipped and constants are
        evaluated once (first pass) and kept."""
        from . import kernels
        prev = kernels.storage
        kernels.storage = self.storage
        try:
            self._run_pass(verbose, capture)
        finally:
            kernels.storage = prev

    def _run_pass(self, verbose, capture):
        from . import kernels
        from .device import is_device
        G = self.ienet.G
        p = self.ienet.ie.plugins
        plan = self._plan if self._plan is not None else self.build_plan()
        concat_bufs = {}
        group_done = set()
        from . import device as _dev
        arena = _dev.current_arena()
        live = _Liveness(self, arena) if (capture and arena is not None and self.reuse_buffers) else None
        self._live = live
        for task in self.task_list:
            node = G.nodes[task]
            step = plan[task]
            if step['skip']:
                continue
            if task in group_done:
                continue
            if step.get('group') and self.kernel_type not in ('fp32', 'tf32x3', 'tf32', 'safe', 'f16x2') and \
                    kernels.default_math == kernels._cabi.MATH_AUTO and self.expected_result is None and not self.pickle_node_args:
                if self._run_group(step['group'], concat_bufs):
                    group_done.update(step['group'])
                    if live is not None:
                        for m in step['group']:
                            live.consumed(m)
                    continue
            node_type = node['type']
            if node_type not in p.plugins:
                print('ERROR: Operation \'{}\' (node={}) is not supported.'.format(node_type, node['name']))
                sys.exit(-1)
            if step.get('const_done'):
                continue
            inputs = self.prepare_inputs_for_task(task) if 'input' in node else {}
            if 'pre_pool' in step['ops']:
                inputs[0] = self.prepare_inputs_for_task(step['ops']['pre_pool'])[0]     # the folded MaxPool's own input
            fused = {}
            for key, val in step['ops'].items():
                if key in ('bias', 'scale', 'shift'):
                    cnode = G.nodes[val]
                    if 'data' not in cnode['output'][0]:
                        # eager fused mode: a Parameter scheduled ahead of the Const it folded (`data/mean`) -- evaluate it now
                        cnode['output'][0]['data'] = p.plugins['Const'].compute(cnode, {}, kernel_type=self.kernel_type, debug=False)[0]
                    fused[key] = cnode['output'][0]['data']
                elif key == 'hl_out' and (self.expected_result is not None or self.pickle_node_args or verbose):
                    continue                     # debug modes read every node output on the host: keep it FP32
                else:
                    fused[key] = val
            if step['out_slot'] is not None:
                cid, coff, c = step['out_slot']
                if cid not in concat_bufs:
                    n_, c_, h_, w_ = G.nodes[cid]['output'][common_def.first_output_port(G.nodes[cid])]['dims']
                    concat_bufs[cid] = kernels.new_nhwc(n_, c_, h_, w_)
                fused['out'] = kernels.channel_slice(concat_bufs[cid], coff, c)
            if node_type == 'Concat' and 'concat_slots' in step and self.fuse:
                if task not in concat_bufs:
                    n_, c_, h_, w_ = node['output'][common_def.first_output_port(node)]['dims']
                    concat_bufs[task] = kernels.new_nhwc(n_, c_, h_, w_)
                buf = concat_bufs[task]
                for src, (coff, c) in step['concat_slots'].items():
                    src_head = next((h for h in plan if plan[h]['store_as'] == src and not plan[h]['skip']), None)
                    if src_head is None or plan[src_head]['out_slot'] is None or plan[src_head]['out_slot'][0] != task:
                        port = [conn for conn in (G.edges[(src, task)]['connection'],)][0][1]
                        kernels.copy_channels(kernels.as_nhwc(G.nodes[src]['output'][port]['data']),
                                              kernels.channel_slice(buf, coff, c))
                fused['inplace'] = buf
            if capture and node_type == 'Result':
                x = inputs[0]
                self._static_out[node['name']] = kernels.as_plain(x) if is_device(x) else x
                if live is not None:
                    live.pin(x)
                    live.pin(self._static_out[node['name']])
                continue
            if verbose:
                print('{}, {}, {}, '.format(task, node_type, node['name']), end=' ', flush=True)
            if task in self.pickle_node_args and not capture:
                with open('node_args_{}.pickle'.format(task), 'wb') as f:
                    host_inputs = {k: (np.asarray(v) if is_device(v) else v) for k, v in inputs.items()}
                    pickle.dump(({k: v for k, v in node.items() if k not in ('output', 'const')}, host_inputs), file=f)
            stime = time.time()
            ev = None
            if self._step_events is not None and node_type != 'Const':
                import torch
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if fused:
                res = p.plugins[node_type].compute(node, inputs, kernel_type=self.kernel_type, debug=False, fused=fused)
            else:
                res = p.plugins[node_type].compute(node, inputs, kernel_type=self.kernel_type, debug=False)
            if ev is not None:
                ev[1].record()
                self._step_events.append((task, ev[0], ev[1]))
            if verbose:
                import torch
                torch.cuda.current_stream().synchronize()
                etime = time.time()
                print(etime - stime)
                self.last_node_seconds[node['name']] = etime - stime
            if self.expected_result is not None and node['name'] in self.expected_result and len(res) > 0:
                out_data = np.asarray(next(iter(res.values())))
                gt = np.asarray(self.expected_result[node['name']][2]).astype(out_data.dtype)
                ok = out_data.shape == gt.shape and np.allclose(out_data, gt, rtol=1e-4, atol=1e-5)
                print('{}{} : {} / {}\x1b[37m'.format('\x1b[32m' if ok else '\x1b[31m', node['name'], out_data.shape, gt.shape))
            if len(res) > 0:
                target = G.nodes[step['store_as']]
                for port_id, data in res.items():
                    tport = port_id if step['store_as'] == task else common_def.first_output_port(target)
                    target['output'][tport]['data'] = data
                    if live is not None:
                        live.stored(step['store_as'], data)
            if live is not None:
                live.consumed(task)
        if live is not None:
            live.sweep()

    def _run_group(self, members, concat_bufs):
        """Sibling 1x1 convolutions as one multi-output contraction.  False -> run the members one by one."""
        from . import kernels
        from .device import is_device
        G = self.ienet.G
        plan = self._plan
        first = G.nodes[members[0]]
        x = self.prepare_inputs_for_task(members[0])[0]
        if not is_device(x) or x.layout != 'nhwc':
            return False
        # what b200ov_conv2d_multi (f16x2 only) accepts -- otherwise the members run one by one and each picks its own
        # kernel (conv_f16x2.cu: f16x2_eligible)
        if x.ptr % 16 != 0 or x.ld % (8 if x.st == 'f16' else 4) != 0 or x.shape[1] % 8 != 0:
            return False
        specs, act = [], plan[members[0]]['ops'].get('act')
        for m in members:
            node, st = G.nodes[m], plan[m]
            ins = self.prepare_inputs_for_task(m)
            bias = G.nodes[st['ops']['bias']]['output'][0]['data'] if 'bias' in st['ops'] else None
            out = None
            if st['out_slot'] is not None:
                cid, coff, c = st['out_slot']
                if cid not in concat_bufs:
                    n_, c_, h_, w_ = G.nodes[cid]['output'][common_def.first_output_port(G.nodes[cid])]['dims']
                    concat_bufs[cid] = kernels.new_nhwc(n_, c_, h_, w_)
                out = kernels.channel_slice(concat_bufs[cid], coff, c)
            specs.append((ins[1], bias, out, bool(st['ops'].get('hl_out')) and out is None))
        ev = None
        if self._step_events is not None:
            import torch
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        outs = kernels.conv1x1_group(x, specs, act=act)
        if ev is not None:
            ev[1].record()
            self._step_events.append((members[0], ev[0], ev[1]))
        for m, y in zip(members, outs):
            target = G.nodes[plan[m]['store_as']]
            target['output'][common_def.first_output_port(target)]['data'] = y
            if getattr(self, '_live', None) is not None:
                self._live.stored(plan[m]['store_as'], y)
        return True

    def run_tasks(self, verbose: bool = False):
        """Eager pass: every scheduled node through its plugin (inference_engine.py:259-292)."""
        self._run(verbose=verbose, capture=False)

    def _mark_constants(self):
        """Evaluate input-independent nodes once (weights upload, SSD prior-box branch) and keep them."""
        G = self.ienet.G
        plan = self._plan
        const_nodes = set()
        for task in self.task_list:
            t = G.nodes[task]['type']
            if t == 'Const' or t == 'ShapeOf':
                const_nodes.add(task)
            elif t not in ('Parameter', 'Result') and all(p in const_nodes for p in G.pred[task]) and G.in_degree(task) > 0:
                const_nodes.add(task)
        return const_nodes

    # OpenVINO IE compatible API - Run inference
    def infer(self, inputs: dict, verbose: bool = False) -> dict:
        dev = self._ensure_device()
        import torch
        G = self.ienet.G
        if self._plan is None:
            self.build_plan()
        graph_mode = self.use_graph and not verbose and not self.pickle_node_args and self.expected_result is None
        self._user_inputs = dict(inputs)
        if not graph_mode or self._graph is None:
            for node_name, val in inputs.items():        # inference_engine.py:300-303
                for node in G.nodes:
                    if G.nodes[node]['name'] == node_name:
                        G.nodes[node]['param'] = val
        if verbose:
            print('# node_id node_name time (sec)')
        stime = time.time()
        from . import kernels
        with torch.cuda.stream(self.stream):
            kernels.status_reset()
            if graph_mode:
                res = self._infer_graph()
            else:
                self.run_tasks(verbose)
                kernels.status_fetch()
                self.stream.synchronize()
                res = {G.nodes[n]['name']: G.nodes[n]['result'] for n, _ in self.ienet.find_node_by_type('Result')}
            if kernels.status_value() != 0:
                res = self._infer_full_range(inputs, verbose)
        etime = time.time()
        if verbose:
            print('@TOTAL_TIME,', etime - stime)
        return res

    # ---- CUDA-graph replay path ----------------------------------------------------------------------
    # Host inputs keep their native element type across PCIe (uint8 / int8 / float16 / float32, `device.RawInput`) and are
    # widened by the layout kernel, which is part of the captured launch sequence.  One graph is captured per input-dtype
    # signature ("key"), lazily; all of them share the arena and the weights.
    def _param_nodes(self):
        G = self.ienet.G
        return [G.nodes[n] for n, _ in self.ienet.find_node_by_type('Parameter')]

    def _ensure_static_in(self):
        if not self._static_in:
            for node in self._param_nodes():
                shape = tuple(node['data']['shape'])
                self._static_in[node['name']] = {'node': node, 'shape': shape, 'n': int(np.prod(shape)), 'bufs': {}}
        return self._static_in

    def _static_buf(self, name, np_dtype):
        """{'host': pinned tensor, 'dev': RawInput} staging pair of input `name` for element type `np_dtype`."""
        import torch
        from . import device as dev
        self._ensure_device()
        st = self._ensure_static_in()[name]
        dt = np.dtype(np_dtype)
        if dt not in st['bufs']:
            tdt = dev.RawInput.TORCH[dt]
            st['bufs'][dt] = {'host': dev.pinned_empty(st['n'], tdt),
                              'dev': dev.RawInput(torch.empty(st['n'], dtype=tdt, device='cuda'), st['shape'], dt)}
        return st['bufs'][dt]

    def _input_key(self, inputs: dict):
        """Element type each input crosses PCIe in, e.g. (('data', '|u1'),): selects the captured graph."""
        from . import device as dev
        key = []
        for name in self._ensure_static_in():
            if inputs is not None and name in inputs:
                dt = dev.native_input(inputs[name]).dtype
            elif self._graph_key is not None:
                dt = np.dtype(dict(self._graph_key)[name])
            else:
                dt = np.dtype(np.float32)
            key.append((name, np.dtype(dt).str))
        return tuple(key)

    def _prepare_graph(self, key=None):
        """Warm-up (sizes the arena, uploads / packs constants) then capture the launch sequence for input signature `key`."""
        import ctypes as C
        from . import _cabi, kernels
        from . import device as dev
        G = self.ienet.G
        self.load_constants()
        if key is None:
            key = self._input_key(self._user_inputs)
        # 1) constants: evaluated once, outside the arena, then frozen
        consts = self._mark_constants()
        dev.set_arena(None)
        for task in self.task_list:
            if task in consts and not self._plan[task].get('const_done'):
                node = G.nodes[task]
                if self._plan[task]['skip']:
                    continue
                inputs = self.prepare_inputs_for_task(task) if 'input' in node else {}
                res = self.ienet.ie.plugins.plugins[node['type']].compute(node, inputs, kernel_type=self.kernel_type, debug=False)
                for port_id, data in res.items():
                    if isinstance(data, np.ndarray) and data.dtype == np.float32:
                        # folded float constants (SSD prior boxes) become resident device tensors once, here
                        data = kernels.upload(data, keep_host=data.size <= 4096)
                    node['output'][port_id]['data'] = data
                self._plan[task]['const_done'] = True
        # 2) warm-up pass inside the arena (also packs weights: those allocations are persistent)
        if self._arena is None:
            self._arena = dev.Arena()
        self._graph_key = key
        self.stage_inputs(self._user_inputs)
        for name, dt in key:
            self._static_in[name]['node']['param'] = self._static_buf(name, dt)['dev']
        dev.set_arena(self._arena)
        try:
            self._arena.frozen = False
            self._arena.reset()
            self._static_out = {}
            self._run(capture=True)
            self.stream.synchronize()
            # 3) capture
            self._arena.reset()
            self._arena.frozen = True
            self._static_out = {}
            s = C.c_void_p(self.stream.cuda_stream)
            launches0 = _cabi.launch_count
            _cabi.call('b200ov_graph_begin', s)
            try:
                self._run(capture=True)
            finally:
                g = C.c_void_p(0)
                _cabi.call('b200ov_graph_end', s, C.byref(g))
            self._graph_launches = _cabi.launch_count - launches0
            self._captured[key] = {'graph': g, 'static_out': self._static_out, 'launches': self._graph_launches}
            self._graph = g
        finally:
            dev.set_arena(None)
        if not getattr(self, '_out_host', None):
            self._out_host = {name: dev.pinned_empty(arr.size) for name, arr in self._static_out.items()}

    def _select_graph(self, inputs: dict):
        """Make the graph captured for the dtype signature of `inputs` current (capturing it first if needed)."""
        key = self._input_key(inputs)
        if key not in self._captured:
            self._user_inputs = dict(inputs) if inputs else self._user_inputs
            self._prepare_graph(key)
        cap = self._captured[key]
        self._graph, self._graph_key, self._static_out, self._graph_launches = cap['graph'], key, cap['static_out'], cap['launches']
        for name, dt in key:
            self._static_in[name]['node']['param'] = self._static_buf(name, dt)['dev']
        return key

    def stage_inputs(self, inputs: dict = None):
        """Copy host inputs (default: the arrays given to the last infer()) into the static device buffers of their
        element type.  An array that already IS the pinned staging buffer (`input_buffer()`) is not copied on the host."""
        from . import device as dev
        for name, st in self._ensure_static_in().items():
            val = inputs[name] if inputs is not None and name in inputs else None
            if val is None:
                continue
            a = dev.native_input(val)
            buf = self._static_buf(name, a.dtype)
            staging = buf['host'].numpy()
            assert a.size == staging.size, 'input {} has {} elements, network expects {}'.format(name, a.size, staging.size)
            if a.__array_interface__['data'][0] != staging.__array_interface__['data'][0]:
                staging[:] = a.reshape(-1)             # ordinary (pageable) user array: one host copy into pinned memory
            buf['dev'].t.copy_(buf['host'], non_blocking=True)

    def input_buffer(self, name: str, dtype=np.float32):
        """Pinned host ndarray (IR shape, element type `dtype`: float32 / float16 / uint8 / int8) for input `name`.
        Filling it in place and passing it to `infer()` skips the pageable->pinned staging copy: the H2D DMA reads it
        directly."""
        st = self._ensure_static_in()[name]
        return self._static_buf(name, dtype)['host'].numpy().reshape(st['shape'])

    def replay(self):
        """Launch the current captured graph on self.stream (inputs must already be staged)."""
        import ctypes as C
        from . import _cabi
        _cabi.call('b200ov_graph_launch', self._graph, C.c_void_p(self.stream.cuda_stream))

    # ---- asynchronous requests (the reference accepts `num_requests` and ignores it, inference_engine.py:86) ----
    NUM_REQUESTS = 2

    def _ensure_requests(self):
        import torch
        from . import device as dev
        if not getattr(self, '_requests', None):
            self._copy_stream = torch.cuda.Stream()
            self._requests = [{'host': {}, 'dev': {}, 'out_host': {}, 'busy': False, 'inputs': None, 'key': None,
                               'h2d_done': torch.cuda.Event(), 'done': torch.cuda.Event(),
                               'status': dev.pinned_empty(1, torch.int32, zero=True)} for _ in range(self.NUM_REQUESTS)]
            self._next_request = 0
        return self._requests

    def _request_buf(self, slot, name, np_dtype):
        import torch
        from . import device as dev
        rq = self._ensure_requests()[slot]
        st = self._ensure_static_in()[name]
        k = (name, np.dtype(np_dtype).str)
        if k not in rq['host']:
            tdt = dev.RawInput.TORCH[np.dtype(np_dtype)]
            rq['host'][k] = dev.pinned_empty(st['n'], tdt)
            rq['dev'][k] = torch.empty(st['n'], dtype=tdt, device='cuda')
        return rq['host'][k], rq['dev'][k]

    def next_slot(self) -> int:
        """The request slot the next `start_async(inputs)` (without an explicit `slot`) will use."""
        self._ensure_device()
        self._ensure_requests()
        return self._next_request

    def request_buffer(self, slot: int, name: str, dtype=np.float32):
        """Pinned host ndarray (IR shape, element type `dtype`) owned by request slot `slot` for input `name`.  Fill it
        and pass it to `start_async(..., slot=slot)`: the H2D DMA then reads it in place (no host copy)."""
        self._ensure_device()
        st = self._ensure_static_in()[name]
        return self._request_buf(slot, name, dtype)[0].numpy().reshape(st['shape'])

    def start_async(self, inputs: dict, slot: int = None) -> int:
        """Queue one inference and return its request slot.  NUM_REQUESTS slots: the H2D copy of one request runs on a
        copy stream while the previous one computes, so a caller that keeps two requests in flight (start_async(i+1)
        before wait(i)) hides the PCIe transfer behind the kernels.  `inputs` maps input names to host arrays in their
        native element type.  `slot` pins the request to a slot (default: round robin, see `next_slot()`); an input
        that is the slot's own `request_buffer(slot, name, dtype)` is transferred without a host copy, any other array
        is first copied into it."""
        import torch
        from . import kernels
        from . import device as dev
        self._ensure_device()
        if self._plan is None:
            self.build_plan()
        self._ensure_requests()
        with torch.cuda.stream(self.stream):
            key = self._select_graph(inputs)
        cap = self._captured[key]
        if slot is None:
            slot = self._next_request
        assert 0 <= slot < self.NUM_REQUESTS
        self._next_request = (slot + 1) % self.NUM_REQUESTS
        rq = self._requests[slot]
        if rq['busy']:
            rq['done'].synchronize()               # the slot's previous request must have drained
        rq['inputs'], rq['key'] = dict(inputs), key
        for name, arr in cap['static_out'].items():
            if name not in rq['out_host']:
                rq['out_host'][name] = dev.pinned_empty(arr.size)
        with torch.cuda.stream(self._copy_stream):
            for name, dt in key:
                if name not in inputs:
                    continue
                host, devbuf = self._request_buf(slot, name, dt)
                staging = host.numpy()
                a = dev.native_input(inputs[name])
                assert a.size == staging.size, 'input {} has {} elements, network expects {}'.format(name, a.size, staging.size)
                if a.__array_interface__['data'][0] != staging.__array_interface__['data'][0]:
                    staging[:] = a.reshape(-1)
                devbuf.copy_(host, non_blocking=True)
            rq['h2d_done'].record(self._copy_stream)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(rq['h2d_done'])
            kernels.status_reset()
            for name, dt in key:
                if name in inputs:                     # D2D into the graph's input buffer
                    self._static_buf(name, dt)['dev'].t.copy_(self._request_buf(slot, name, dt)[1], non_blocking=True)
            self.replay()
            for name, arr in cap['static_out'].items():
                rq['out_host'][name][:arr.size].copy_(arr.t[:arr.size], non_blocking=True)
            kernels.status_fetch(rq['status'])
            rq['done'].record(self.stream)
        rq['busy'] = True
        return slot

    def wait(self, slot: int) -> dict:
        """Block until request `slot` has finished and return {result_name: ndarray}."""
        rq = self._requests[slot]
        rq['done'].synchronize()
        rq['busy'] = False
        if int(rq['status'][0]) != 0:
            import torch
            with torch.cuda.stream(self.stream):
                return self._infer_full_range(rq['inputs'])
        outs = self._captured[rq['key']]['static_out']
        return {name: rq['out_host'][name][:arr.size].numpy().reshape(arr.shape).copy() for name, arr in outs.items()}

    def _infer_full_range(self, inputs: dict, verbose: bool = False):
        """The f16x2 contractions saw a non-finite output: an operand left the FP16 range (|v| > 65504) or the
        data holds inf / NaN.  Repeat this inference eagerly with the FP32-range kernels (3xTF32 / FFMA), whose
        results follow the reference for any finite FP32 input.  Rare by construction; counted in
        `self.range_fallbacks`."""
        from . import _cabi, kernels
        from . import device as dev
        G = self.ienet.G
        self.range_fallbacks = getattr(self, 'range_fallbacks', 0) + 1
        saved_params = {}
        for node_name, val in inputs.items():
            for node in G.nodes:
                if G.nodes[node]['name'] == node_name:
                    saved_params[node] = G.nodes[node].get('param')
                    G.nodes[node]['param'] = val
        saved_math, saved_storage = kernels.default_math, self.storage
        kernels.default_math = _cabi.MATH_SAFE
        self.storage = 'f32'                 # a value beyond FP16 cannot be STORED as FP16 either
        dev.set_arena(None)
        try:
            self._run(verbose=verbose, capture=False)
            self.stream.synchronize()
        finally:
            kernels.default_math = saved_math
            self.storage = saved_storage
            for node, val in saved_params.items():
                G.nodes[node]['param'] = val
        return {G.nodes[n]['name']: G.nodes[n]['result'] for n, _ in self.ienet.find_node_by_type('Result')}

    def fetch_outputs(self):
        from . import kernels
        res = {}
        for name, arr in self._static_out.items():
            self._out_host[name][:arr.size].copy_(arr.t[:arr.size], non_blocking=True)
        kernels.status_fetch()
        self.stream.synchronize()
        for name, arr in self._static_out.items():
            res[name] = self._out_host[name][:arr.size].numpy().reshape(arr.shape).copy()
        return res

    def _infer_graph(self):
        self._select_graph(self._user_inputs)
        self.stage_inputs(self._user_inputs)
        self.replay()
        return self.fetch_outputs()

    def load_constants(self):
        """Upload every float Const into ONE resident HBM buffer (the weight arena) with a single H2D copy;
        each Const node's DeviceArray is a view into it.  Returns the flat torch tensor (what
        `distributed.broadcast_weights` sends from rank 0)."""
        import torch
        from .device import DeviceArray
        dev = self._ensure_device()
        if self._weights is not None:
            return self._weights
        G = self.ienet.G
        items, total = [], 0
        for node_id, _ in self.ienet.find_node_by_type('Const'):
            node = G.nodes[node_id]
            if common_def.type_convert_tbl[node['data']['element_type']] is not np.float32:
                continue
            n = int(node['const']['data'].size)
            items.append((node, total, n))
            total += (n + 63) // 64 * 64
        host = dev.pinned_empty(total, zero=True)
        hv = host.numpy()
        for node, off, n in items:
            hv[off:off + n] = node['const']['data']
        with torch.cuda.stream(self.stream):
            flat = torch.empty(max(total, 1), dtype=torch.float32, device='cuda')
            flat.copy_(host, non_blocking=True)
            self.stream.synchronize()
        for node, off, n in items:
            arr = DeviceArray(flat[off:off + max(n, 1)], node['data']['shape'], 'plain')
            if n <= 4096:
                arr.cache['host'] = np.array(node['const']['data'], dtype=np.float32).reshape(node['data']['shape'])
            node['const']['device'] = arr
        self._weights = flat
        return flat

    def broadcast_constants(self, src: int = 0):
        """Multi-GPU load step: make rank `src`'s weight arena the replica every rank computes with.
        Synchronises both sides of the NCCL broadcast and rebuilds everything derived from the constants on this
        rank -- the small host mirrors and the packed / split weight forms -- so nothing stale survives when the
        ranks did not load identical `.bin` files.  Must run before the first inference."""
        from . import distributed
        flat = self.load_constants()
        rank, size, _ = distributed.world()
        distributed.broadcast_weights(flat, src=src)
        if size > 1 and rank != src:
            G = self.ienet.G
            for node_id, _name in self.ienet.find_node_by_type('Const'):
                arr = G.nodes[node_id]['const'].get('device')
                if arr is None:
                    continue
                had_host = 'host' in arr.cache
                arr.cache.clear()
                if had_host:
                    arr.cache['host'] = arr.numpy()
        return flat

    def profile_steps(self, inputs: dict, iters: int = 3):
        """Per-step device time of the fused plan run eagerly (CUDA events around every plugin call on
        self.stream).  Returns [{'id', 'type', 'name', 'ms'}] averaged over `iters` passes."""
        import torch
        from . import device as dev
        self._ensure_device()
        if self._plan is None:
            self.build_plan()
        G = self.ienet.G
        self._user_inputs = dict(inputs)
        acc = {}
        with torch.cuda.stream(self.stream):
            if self.use_graph:
                self._select_graph(inputs)
                self.stage_inputs(inputs)
            else:
                for name, val in inputs.items():
                    for node in G.nodes:
                        if G.nodes[node]['name'] == name:
                            G.nodes[node]['param'] = val
            self.load_constants()        # weights must never come out of the per-inference arena
            arena = self._arena if self._arena is not None else dev.Arena()
            frozen = arena.frozen
            arena.frozen = False
            dev.set_arena(arena)
            try:
                for it in range(iters + 1):
                    arena.reset()
                    self._step_events = []
                    # keep the GPU busy while the host queues the whole pass, so the events around each step
                    # measure kernel time and not the host's launch gaps
                    torch.cuda._sleep(40000000)
                    self._run(capture=self.use_graph)
                    self.stream.synchronize()
                    if it > 0:
                        for task, e0, e1 in self._step_events:
                            acc[task] = acc.get(task, 0.0) + e0.elapsed_time(e1)
                    self._step_events = None
            finally:
                dev.set_arena(None)
                arena.frozen = frozen
                self._step_events = None
        return [{'id': t, 'type': G.nodes[t]['type'], 'name': G.nodes[t]['name'], 'ms': ms / iters} for t, ms in acc.items()]

    def kernels_per_inference(self):
        """Number of libb200ov kernel launches captured in the graph (0 before the first inference)."""
        return getattr(self, '_graph_launches', 0)
