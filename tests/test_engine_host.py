"""Host logic of the engine (CPU only): IR parsing, scheduling, re-batching, fusion plan, plugin
discovery and the host-side glue plugins."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, REPO

MODELS = ['mnist', 'mnist_bn', 'googlenet-v1', 'ssd_mobilenet_v1_coco']


@pytest.fixture(scope='module')
def ie():
    from pyopenvino_b200.inference_engine import IECore
    return IECore()


def test_plugin_discovery_file_name_is_type(ie):
    have = set(ie.plugins.plugins)
    for t in ('Convolution', 'GroupConvolution', 'MatMul', 'MaxPool', 'AvgPool', 'Add', 'Multiply', 'ReLU', 'Clamp', 'SoftMax',
              'Const', 'Parameter', 'Result', 'LRN', 'Concat', 'Sigmoid', 'Transpose', 'Reshape', 'Unsqueeze'):
        assert t in have
        assert callable(ie.plugins.plugins[t].compute) and callable(ie.plugins.plugins[t].name)


@pytest.mark.parametrize('model', MODELS)
def test_ir_graph_matches_oracle_parse(ie, model, model_dir):
    from oracle import ref_engine
    path = os.path.join(model_dir, model + '.xml')
    net = ie.read_network(path, path[:-4] + '.bin')
    ref = ref_engine.RefNetwork(path)
    assert set(net.G.nodes) == set(ref.nodes)
    assert ie.check_nodes(net.G) == set()
    for nid, rn in ref.nodes.items():
        n = net.G.nodes[nid]
        assert n['type'] == rn['type'] and n['name'] == rn['name']
        assert n.get('input') == rn.get('input') and {p: {k: v for k, v in d.items() if k != 'data'} for p, d in n.get('output', {}).items()} == rn.get('output', {})
        if n['type'] == 'Const':
            assert np.array_equal(np.asarray(n['const']['data']), np.asarray(rn['const']))
    assert sorted(net.edges) == sorted(ref.edges)
    exe = ie.load_network(net, fuse=False, use_graph=False)
    assert sorted(exe.task_list) == sorted(ref.order)
    pos = {n: i for i, n in enumerate(exe.task_list)}
    for fl, fp, tl, tp in net.edges:
        assert pos[fl] < pos[tl]
    assert net.inputs[0]['type'] == 'Parameter' and net.outputs[0]['type'] == 'Result'


def test_missing_model_and_bad_xml_raise(ie, tmp_path):
    with pytest.raises(Exception):
        ie.read_network(str(tmp_path / 'nope.xml'), 'x')
    (tmp_path / 'bad.xml').write_text('<notnet/>')
    (tmp_path / 'bad.bin').write_bytes(b'')
    with pytest.raises(Exception):
        ie.read_network(str(tmp_path / 'bad.xml'), 'x')


@pytest.mark.parametrize('model,batch', [('mnist_bn', 1024), ('googlenet-v1', 8), ('ssd_mobilenet_v1_coco', 3)])
def test_rebatching_rewrites_every_dynamic_port(ie, model, batch, model_dir):
    path = os.path.join(model_dir, model + '.xml')
    net = ie.read_network(path, None)
    before = {n: json.dumps(net.G.nodes[n].get('output', {}), default=str) for n in net.G.nodes}
    net.set_batch_size(batch)
    G = net.G
    assert net.inputs[0]['data']['shape'][0] == batch
    for fl, fp, tl, tp in net.edges:
        assert G.nodes[fl]['output'][fp]['dims'] == G.nodes[tl]['input'][tp]['dims']
    for n in G.nodes:
        node = G.nodes[n]
        if node['type'] == 'Const':
            assert json.dumps(node['output'], default=str) == before[n]
    out_dims = net.outputs[0]['input'][0]['dims']
    if model.startswith('ssd'):
        assert out_dims == (1, 1, 100 * batch, 7)
    else:
        assert out_dims[0] == batch
    net.set_batch_size(1)
    assert all(json.dumps(G.nodes[n].get('output', {}), default=str) == before[n] for n in G.nodes)


def test_fusion_plan_census(ie, model_dir):
    """every Add / ReLU / Clamp / BN Multiply is absorbed by a producer."""
    # GoogLeNet: 86 steps, nine of them 3x3 / stride-1 MaxPools that run inside the pool_proj convolution's producers
    expect = {'mnist': 11, 'mnist_bn': 13, 'googlenet-v1': 77}
    for model, steps in expect.items():
        path = os.path.join(model_dir, model + '.xml')
        net = ie.read_network(path, None)
        exe = ie.load_network(net)
        plan = exe.build_plan()
        G = net.G
        live = [n for n in exe.task_list if not plan[n]['skip'] and G.nodes[n]['type'] not in ('Const', 'Result')]
        types = [G.nodes[n]['type'] for n in live]
        for t in ('Add', 'ReLU', 'Clamp', 'Multiply'):
            assert t not in types, (model, t)
        assert len(live) == steps, (model, len(live))
        folded = [n for n in exe.task_list if plan[n].get('pool_into') is not None]
        assert len(folded) == (9 if model == 'googlenet-v1' else 0)
        for n in folded:
            assert G.nodes[n]['type'] == 'MaxPool' and plan[plan[n]['pool_into']]['ops']['pre_pool'] == n
    # FP16 storage keeps the pools as kernels of their own (half2 max there; the fused producer path is FP32-only)
    net = ie.read_network(os.path.join(model_dir, 'googlenet-v1.xml'), None)
    exe = ie.load_network(net, storage='f16')
    plan = exe.build_plan()
    assert not any(plan[n].get('pool_into') is not None for n in exe.task_list)
    net = ie.read_network(os.path.join(model_dir, 'ssd_mobilenet_v1_coco.xml'), None)
    exe = ie.load_network(net)
    plan = exe.build_plan()
    live = [net.G.nodes[n]['type'] for n in exe.task_list if not plan[n]['skip']]
    assert 'Clamp' not in live and 'Multiply' not in live and 'Add' not in live


def test_host_glue_plugins_vs_reference_vectors(ie):
    ops = np.load(os.path.join(GOLDEN, 'ops.npz'))
    meta = json.loads(str(ops['meta']))
    prec = {np.dtype('float32'): 'FP32', np.dtype('int64'): 'I64'}
    seen = set()
    for i, m in enumerate(meta):
        if m['type'] not in ('ShapeOf', 'StridedSlice', 'PriorBoxClustered'):
            continue
        ins = {p: ops['c{}_in{}'.format(i, p)] for p in m['ports']}
        op = 1 if m['type'] == 'ShapeOf' else len(ins)
        node = {'name': m['tag'], 'type': m['type'], 'data': dict(m['data']),
                'input': {p: {'precision': prec[a.dtype], 'dims': tuple(a.shape)} for p, a in ins.items()},
                'output': {op: {'precision': 'I64' if m['type'] == 'ShapeOf' else 'FP32', 'dims': ()}}}
        if m['type'] == 'ShapeOf':
            ins = {0: np.zeros(ins[0].shape, dtype=np.float32)}
        got = ie.plugins.plugins[m['type']].compute(node, ins, kernel_type='numpy')[op]
        want = ops['c{}_out_numpy'.format(i)]
        assert got.dtype == want.dtype and np.array_equal(got, want), m['tag']
        seen.add(m['type'])
    assert seen == {'ShapeOf', 'StridedSlice', 'PriorBoxClustered'}


def test_shape_rules_match_oracle():
    from oracle import ref_ops
    from pyopenvino_b200 import common_def
    from pyopenvino_b200.op_plugins.Reshape import resolve_shape
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = rng.integers(1, 40, 2)
        k = int(rng.integers(1, 8))
        s = int(rng.integers(1, 4))
        pb, pe = rng.integers(0, 3, 2), rng.integers(0, 3, 2)
        if h + pb[0] + pe[0] < k or w + pb[1] + pe[1] < k or h < k or w < k:
            continue
        for rounding in ('floor', 'ceil'):
            for ap in ('explicit', 'valid', 'same_upper'):
                for same_ceil in (True, False):
                    a = common_def.spatial_output_shape((h, w), (k, k), (s, s), pb, pe, rounding, ap, same_ceil)
                    b = ref_ops.out_hw((h, w), (k, k), (s, s), pb, pe, rounding, ap, same_ceil)
                    assert tuple(a) == tuple(b)
    assert resolve_shape((1, 5, 7, 12), [0, -1]) == (1, 420)
    assert resolve_shape((3, 3, 3, 64), [-1, 576]) == (3, 576)
    assert resolve_shape((2, 3, 3, 12), [0, -1, 1, 4]) == (2, 27, 1, 4)
    with pytest.raises(AssertionError):
        resolve_shape((2, 3), [-1, -1])
    # a target without -1 must keep the element count (the reference's ndarray.reshape raises, Reshape.py:44):
    # a hard-coded batch-1 target on a re-batched network must not silently drop images
    with pytest.raises(ValueError):
        resolve_shape((4, 3, 3, 64), [1, 576])
    with pytest.raises(ValueError):
        np.zeros((4, 3, 3, 64)).reshape((1, 576))
    assert resolve_shape((4, 3, 3, 64), [4, 576]) == (4, 576)


def test_no_product_module_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pyopenvino_b200/ may reference it."""
    pkg = os.path.join(REPO, 'pyopenvino_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(root, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'ref_ops' not in text, f


def test_missing_library_fails_loudly(monkeypatch):
    from pyopenvino_b200 import _cabi
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', '/nonexistent/libb200ov.so')
    with pytest.raises(_cabi.B200ovError):
        _cabi.load()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pyopenvino_b200 import _cabi, device
    with pytest.raises(_cabi.B200ovError):
        device.init()


def test_operand_form_edge_census(ie, model_dir, monkeypatch):
    """tensors whose every reader is a contraction are planned in the (hi, lo) operand form -- the *_reduce
    outputs of GoogLeNet, depthwise -> pointwise and the 1x1 -> 3x3/s2 pairs of SSD, conv -> conv of mnist_bn; none when a
    MaxPool / Concat / Result reads the tensor, under FP16 storage, or with B200OV_NO_HL=1."""
    expect = {'mnist': 0, 'mnist_bn': 2, 'googlenet-v1': 19, 'ssd_mobilenet_v1_coco': 22}
    for model, count in expect.items():
        path = os.path.join(REPO, 'models', 'mnist.xml') if model == 'mnist' else os.path.join(model_dir, model + '.xml')
        net = ie.read_network(path, None)
        exe = ie.load_network(net)
        plan = exe.build_plan()
        G = net.G
        heads = [n for n in exe.task_list if plan[n]['ops'].get('hl_out')]
        assert len(heads) == count, (model, len(heads))
        for n in heads:
            assert G.nodes[n]['type'] in ('Convolution', 'GroupConvolution') and plan[n]['out_slot'] is None
            for r in G.successors(plan[n]['store_as']):
                assert G.nodes[r]['type'] == 'Convolution' and 'pre_pool' not in plan[r]['ops']
    net = ie.read_network(os.path.join(model_dir, 'googlenet-v1.xml'), None)
    exe = ie.load_network(net, storage='f16')
    assert not any(st['ops'].get('hl_out') for st in exe.build_plan().values())
    monkeypatch.setenv('B200OV_NO_HL', '1')
    net = ie.read_network(os.path.join(model_dir, 'googlenet-v1.xml'), None)
    exe = ie.load_network(net)
    assert not any(st['ops'].get('hl_out') for st in exe.build_plan().values())
