"""GPU parity, whole networks: the reference-facing API (IECore / read_network / load_network /
infer with host arrays) against the reference's golden outputs and the oracle engine."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, REPO, close

pytestmark = pytest.mark.gpu


def _load(model_dir, model, batch=None, fuse=True, use_graph=True):
    from pyopenvino_b200.inference_engine import IECore
    ie = IECore()
    path = os.path.join(REPO, 'models', 'mnist.xml') if model == 'mnist' else os.path.join(model_dir, model + '.xml')
    net = ie.read_network(path, path[:-4] + '.bin')
    exe = ie.load_network(net, 'B200', batch_size=batch, fuse=fuse, use_graph=use_graph)
    return net, exe


@pytest.mark.parametrize('fuse,use_graph', [(False, False), (True, False), (True, True)])
def test_mnist_known_answer(model_dir, fuse, use_graph):
    """README.md:69-72 / integrity_test.py:57 on resources/mnist2.png, real weights."""
    g = np.load(os.path.join(GOLDEN, 'mnist_e2e.npz'))
    net, exe = _load(model_dir, 'mnist', fuse=fuse, use_graph=use_graph)
    exe.kernel_type = 'numpy'
    for _ in range(2):      # second call exercises graph replay
        res = exe.infer({net.inputs[0]['name']: g['input']})
        prob = res[net.outputs[0]['name']]
        assert prob.shape == (1, 10)
        assert list(np.argsort(prob[0])[::-1]) == [2, 0, 1, 7, 8, 6, 3, 4, 5, 9]
        ok, msg = close(prob, g['final_numpy'], rtol=1e-4, atol=1e-7)
        assert ok, msg
    prob7 = exe.infer({net.inputs[0]['name']: g['input7']})[net.outputs[0]['name']]
    ok, msg = close(prob7, g['final7_special'], rtol=1e-4, atol=1e-7)
    assert ok, msg


def test_mnist_every_node_vs_reference(model_dir):
    """Unfused eager mode materialises every node output like the reference does
    (inference_engine.py:290-292); compare all 33 against the reference's own feature maps."""
    g = np.load(os.path.join(GOLDEN, 'mnist_e2e.npz'))
    names = json.loads(str(g['node_names']))
    net, exe = _load(model_dir, 'mnist', fuse=False, use_graph=False)
    exe.infer({net.inputs[0]['name']: g['input']})
    got = {}
    for nid in net.G.nodes:
        n = net.G.nodes[nid]
        if 'output' in n:
            p = next(iter(n['output']))
            if 'data' in n['output'][p]:
                got[n['name']] = np.asarray(n['output'][p]['data'])
    # activations reach ~1e3 on raw 0..255 pixels: tolerance is relative to the tensor scale
    for i, nm in enumerate(names):
        want = g['node_{}'.format(i)]
        scale = max(1.0, float(np.abs(want).max()))
        ok, msg = close(got[nm], want, rtol=1e-4, atol=1e-5 * scale)
        assert ok, (nm, msg)


@pytest.mark.parametrize('model', ['mnist_bn', 'googlenet-v1'])
def test_synthetic_models_vs_reference_golden(model_dir, model):
    from tools.synth_bin import synth_input
    g = np.load(os.path.join(GOLDEN, 'models_e2e.npz'))
    x = synth_input(model, batch=2, seed=1)
    net, exe = _load(model_dir, model, batch=2)
    res = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    for img in range(2):
        want = g['{}|special|{}|final'.format(model, img)]
        ok, msg = close(res[img:img + 1], want, rtol=1e-4, atol=1e-6)
        assert ok, (model, img, msg)
        assert np.argmax(res[img]) == np.argmax(want)


@pytest.mark.parametrize('model', ['mnist_bn', 'googlenet-v1', 'ssd_mobilenet_v1_coco'])
def test_every_node_vs_oracle(model_dir, model):
    """Per-node parity with IDENTICAL inputs (SURVEY.md section 8c methodology): every non-Const node of
    the graph is run through our plugin on the oracle's own input feature maps and compared with the
    oracle's output for that node."""
    from oracle import ref_engine
    from pyopenvino_b200.inference_engine import IECore
    from tools.synth_bin import synth_input
    x = synth_input(model, batch=1, seed=1)
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    oracle.infer({oracle.net.inputs[0]['name']: x})
    plugins = IECore().plugins.plugins
    checked = 0
    exact = {'MaxPool', 'ReLU', 'Add', 'Multiply', 'Clamp', 'Concat', 'Reshape', 'Transpose', 'GroupConvolution', 'Unsqueeze',
             'ShapeOf', 'StridedSlice', 'PriorBoxClustered', 'DetectionOutput'}
    for nid in oracle.net.order:
        rn = oracle.net.nodes[nid]
        if rn['type'] in ('Const', 'Parameter', 'Result'):
            continue
        ins = {tp: oracle.net.nodes[fl]['output'][fp]['data'] for fl, fp, tp in oracle.net.pred[nid]}
        port = next(iter(rn['output']))
        node = {'name': rn['name'], 'type': rn['type'], 'data': dict(rn.get('data', {})), 'input': rn['input'],
                'output': {port: {'precision': rn['output'][port]['precision'], 'dims': rn['output'][port]['dims']}}}
        got = np.asarray(plugins[rn['type']].compute(node, ins, kernel_type='numpy')[port])
        want = rn['output'][port]['data']
        if rn['type'] == 'GroupConvolution':
            # default = packed-FMA kernel (tolerance class); kernel_type='exact' = pairwise kernel, bit-identical
            ok, msg = close(got, want, rtol=1e-4, atol=1e-5)
            assert ok, (rn['name'], rn['type'], msg)
            got = np.asarray(plugins[rn['type']].compute(node, ins, kernel_type='exact')[port])
        if rn['type'] in exact:
            assert np.array_equal(got, want), (rn['name'], rn['type'])
        else:
            ok, msg = close(got, want, rtol=1e-4, atol=1e-5)
            assert ok, (rn['name'], rn['type'], msg)
        checked += 1
    assert checked > 20


@pytest.mark.parametrize('model', ['mnist_bn', 'googlenet-v1'])
def test_unfused_eager_end_to_end_vs_oracle(model_dir, model):
    """The reference-like mode (every node through its plugin, outputs materialised) end to end."""
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    x = synth_input(model, batch=1, seed=1)
    oracle = ref_engine.load(os.path.join(model_dir, model + '.xml'), 'special')
    want = next(iter(oracle.infer({oracle.net.inputs[0]['name']: x}).values()))
    net, exe = _load(model_dir, model, fuse=False, use_graph=False)
    got = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    ok, msg = close(got, want, rtol=1e-4, atol=1e-6)
    assert ok, msg
    assert np.argmax(got) == np.argmax(want)
    # every node output is materialised on the graph like in the reference (inference_engine.py:290-292)
    for nid in net.G.nodes:
        n = net.G.nodes[nid]
        if 'output' in n:
            assert 'data' in n['output'][next(iter(n['output']))], n['name']


def test_batch_is_stack_of_batch1(model_dir):
    """Batched semantics = B independent batch-1 results, bit for bit."""
    from tools.synth_bin import synth_input
    x = synth_input('mnist_bn', batch=5, seed=3)
    net, exe = _load(model_dir, 'mnist_bn', batch=5)
    full = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    net1, exe1 = _load(model_dir, 'mnist_bn', batch=1)
    for i in range(5):
        one = exe1.infer({net1.inputs[0]['name']: x[i:i + 1]})[net1.outputs[0]['name']]
        ok, msg = close(full[i:i + 1], one, rtol=1e-5, atol=1e-7)
        assert ok, msg


def test_fused_graph_equals_unfused_eager(model_dir):
    from tools.synth_bin import synth_input
    x = synth_input('googlenet-v1', batch=3, seed=9)
    net, exe = _load(model_dir, 'googlenet-v1', batch=3)
    a = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    b = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    assert np.array_equal(a, b)                 # replay is deterministic
    net2, exe2 = _load(model_dir, 'googlenet-v1', batch=3, fuse=False, use_graph=False)
    c = exe2.infer({net2.inputs[0]['name']: x})[net2.outputs[0]['name']]
    ok, msg = close(a, c, rtol=1e-5, atol=1e-7)
    assert ok, msg
    assert exe.kernels_per_inference() > 0


def _records(block):
    """Valid records of one image's (keep_top_k, 7) block (up to the -1 terminator)."""
    stop = np.where(block[:, 0] == -1)[0]
    return block[:int(stop[0])] if len(stop) else block


@pytest.mark.parametrize('fuse,use_graph', [(True, True), (False, False)])
def test_ssd_end_to_end_vs_reference_golden(model_dir, fuse, use_graph):
    """SSD-MobileNet-v1 (synthetic weights) batch 2: box ORDER and class ids identical to the reference's
    own run, scores / coordinates within the FP32 tolerance."""
    from tools.synth_bin import synth_input
    g = np.load(os.path.join(GOLDEN, 'models_e2e.npz'))
    model = 'ssd_mobilenet_v1_coco'
    x = synth_input(model, batch=2, seed=1)
    net, exe = _load(model_dir, model, batch=2, fuse=fuse, use_graph=use_graph)
    res = exe.infer({net.inputs[0]['name']: x})[net.outputs[0]['name']]
    assert res.shape == (1, 1, 200, 7)
    for img in range(2):
        want = _records(g['{}|special|{}|final'.format(model, img)][0, 0])
        got = _records(res[0, 0, img * 100:(img + 1) * 100])
        assert len(want) > 5
        assert got.shape == want.shape, (got.shape, want.shape)
        assert np.array_equal(got[:, 0:2], want[:, 0:2])          # rank + class id, in order
        ok, msg = close(got[:, 2:], want[:, 2:], rtol=1e-4, atol=1e-5)
        assert ok, msg


def test_range_fallback_end_to_end(model_dir):
    """An input far outside the FP16 range raises the f16x2 status word; the executor repeats the inference
    with the FP32-range kernels (3xTF32 / FFMA).  Ordinary inputs never fall back.  (The reference itself
    returns NaN for such an input: its SoftMax has no max shift, SURVEY Appendix A.8, so the comparison is
    against this engine pinned to the full-range kernels; those are checked against the oracle per op.)"""
    from tools.synth_bin import synth_input
    x = synth_input('mnist_bn', batch=2, seed=5)
    net, exe = _load(model_dir, 'mnist_bn', batch=2)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    exe.infer({name: x})
    assert getattr(exe, 'range_fallbacks', 0) == 0
    big = x.copy()
    big[:, :, 10:14, 10:14] = 2.0e5         # first conv is C_in = 1 (FFMA); its outputs (~1e5) overflow FP16 in conv 2
    got = exe.infer({name: big})[out]
    assert exe.range_fallbacks == 1
    net2, exe2 = _load(model_dir, 'mnist_bn', batch=2, use_graph=False)
    exe2.kernel_type = 'safe'
    want = exe2.infer({name: big})[out]
    assert np.all(np.isfinite(got)) and np.all(np.isfinite(want))
    assert np.array_equal(got, want)
    again = exe.infer({name: x})[out]       # the graph path is intact afterwards
    assert exe.range_fallbacks == 1
    assert np.all(np.isfinite(again))


def test_async_requests_match_sync_infer(model_dir):
    """start_async / wait with two requests in flight returns what infer() returns for the same inputs."""
    from tools.synth_bin import synth_input
    net, exe = _load(model_dir, 'mnist_bn', batch=4)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    xs = [synth_input('mnist_bn', batch=4, seed=20 + i) for i in range(5)]
    want = [exe.infer({name: x})[out] for x in xs]
    got, pending = [], None
    for x in xs:
        slot = exe.start_async({name: x})
        if pending is not None:
            got.append(exe.wait(pending)[out])
        pending = slot
    got.append(exe.wait(pending)[out])
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


# ---- BASELINE.json full-size configurations: size-independent properties -------------------------------------------

def test_googlenet_batch256_rows_are_batch1_results(model_dir):
    """configs[2] at the bench size (256 images per GPU): the batched result is the stack of the batch-1 results
    (SURVEY.md section 0.4), replay is deterministic, every row is a probability vector, and the rows checked against
    the oracle engine keep its argmax."""
    from oracle import ref_engine
    from tools.synth_bin import synth_input
    x = synth_input('googlenet-v1', batch=256, seed=31)
    net, exe = _load(model_dir, 'googlenet-v1', batch=256)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    full = exe.infer({name: x})[out]
    assert full.shape == (256, 1000)
    again = exe.infer({name: x})[out]
    bad_rows = np.unique(np.nonzero(full != again)[0])
    assert bad_rows.size == 0, 'replay 1 and replay 2 differ in rows {} (max |d| {:.3g})'.format(
        bad_rows[:16].tolist(), float(np.abs(full - again).max()))
    assert np.all(np.isfinite(full)) and np.all(full >= 0)
    assert np.allclose(full.sum(axis=1), 1.0, atol=1e-5)
    net1, exe1 = _load(model_dir, 'googlenet-v1', batch=1)
    oracle = ref_engine.load(os.path.join(model_dir, 'googlenet-v1.xml'), 'special')
    for i in (0, 97, 255):
        one = exe1.infer({net1.inputs[0]['name']: x[i:i + 1]})[net1.outputs[0]['name']]
        ok, msg = close(full[i:i + 1], one, rtol=1e-5, atol=1e-8)
        assert ok, (i, msg)
        want = oracle.infer({name: x[i:i + 1]})[out]
        ok, msg = close(full[i:i + 1], want, rtol=1e-4, atol=1e-6)
        assert ok, (i, msg)
        assert np.argmax(full[i]) == np.argmax(want)


def test_ssd_batch64_records_are_batch1_records(model_dir):
    """configs[3] at the bench size (64 images per GPU): image i's DetectionOutput record block equals the batch-1
    run on image i -- rank and class id bit for bit, scores and boxes within the FP32 tolerance."""
    from tools.synth_bin import synth_input
    x = synth_input('ssd_mobilenet_v1_coco', batch=64, seed=33)
    net, exe = _load(model_dir, 'ssd_mobilenet_v1_coco', batch=64)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    full = exe.infer({name: x})[out]
    keep = full.shape[2] // 64
    assert np.array_equal(full, exe.infer({name: x})[out])
    net1, exe1 = _load(model_dir, 'ssd_mobilenet_v1_coco', batch=1)
    for i in (0, 21, 63):
        one = exe1.infer({net1.inputs[0]['name']: x[i:i + 1]})[net1.outputs[0]['name']]
        blk = full[0, 0, i * keep:(i + 1) * keep]
        ref = one[0, 0, :keep]
        assert np.array_equal(blk[:, 0:2], ref[:, 0:2]), i          # rank + class id, in order
        ok, msg = close(blk[:, 2:], ref[:, 2:], rtol=1e-4, atol=1e-5)
        assert ok, (i, msg)


@pytest.mark.parametrize('model,batch', [('googlenet-v1', 64), ('mnist_bn', 256)])
def test_replay_is_bit_deterministic(model_dir, model, batch):
    """The contraction kernel's roles hand work to each other through named barriers, mbarriers and tensor-memory
    rings; its accumulation order is fixed, so a protocol race would show up as a changing result.  Same batch,
    40 graph replays, one digest."""
    import hashlib
    from tools.synth_bin import synth_input
    x = synth_input(model, batch=batch, seed=7)
    net, exe = _load(model_dir, model, batch=batch)
    name, out = net.inputs[0]['name'], net.outputs[0]['name']
    digests = {hashlib.sha256(np.ascontiguousarray(exe.infer({name: x})[out]).tobytes()).hexdigest() for _ in range(40)}
    assert len(digests) == 1
