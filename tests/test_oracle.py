"""Pins the oracle (oracle/ref_ops.py, oracle/ref_engine.py) to vectors produced by the live
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, close
from oracle import ref_engine, ref_ops

OPS = np.load(os.path.join(GOLDEN, 'ops.npz'))
META = json.loads(str(OPS['meta']))

# ops whose reference arithmetic is numpy's own loops (no BLAS): the restatement must be bit-exact
EXACT = {'GroupConvolution', 'MaxPool', 'AvgPool', 'Add', 'Multiply', 'ReLU', 'Clamp', 'SoftMax', 'Sigmoid', 'LRN',
         'Concat', 'Transpose', 'Reshape', 'Unsqueeze', 'ShapeOf', 'StridedSlice', 'PriorBoxClustered',
         'DetectionOutput'}


def _node(i, m, kt):
    prec = {np.dtype('float32'): 'FP32', np.dtype('int64'): 'I64'}
    ins = {p: OPS['c{}_in{}'.format(i, p)] for p in m['ports']}
    out_port = 1 if m['type'] == 'ShapeOf' else len(ins)
    out_prec = 'I64' if m['type'] == 'ShapeOf' else 'FP32'
    node = {'name': m['tag'], 'type': m['type'], 'data': dict(m['data']),
            'input': {p: {'precision': prec[a.dtype], 'dims': tuple(a.shape)} for p, a in ins.items()},
            'output': {out_port: {'precision': out_prec, 'dims': ()}}}
    return node, ins, out_port


@pytest.mark.parametrize('i', range(len(META)), ids=[m['tag'] for m in META])
def test_op_vectors(i):
    m = META[i]
    for kt in m['kts']:
        node, ins, op = _node(i, m, kt)
        got = ref_engine.run_node(node, ins, kt)[op]
        want = OPS['c{}_out_{}'.format(i, kt)]
        assert got.shape == want.shape and got.dtype == want.dtype
        if m['type'] in EXACT or (m['type'] == 'Convolution' and kt == 'numpy'):
            assert np.array_equal(got, want, equal_nan=True), m['tag']
        else:   # OpenBLAS sgemm behind np.dot / np.matmul: same library here, tolerance elsewhere
            ok, msg = close(got, want, rtol=1e-5, atol=1e-6)
            assert ok, msg


def test_groupconv_vectorised_equals_literal_loops():
    rng = np.random.default_rng(7)
    for (c, hw, s, pb, pe) in [(6, 9, 1, (1, 1), (1, 1)), (4, 10, 2, (0, 0), (1, 1)), (5, 7, 2, (1, 1), (1, 1))]:
        x = rng.standard_normal((1, c, hw, hw)).astype(np.float32)
        w = rng.standard_normal((c, 1, 1, 3, 3)).astype(np.float32)
        a = ref_ops.groupconv_numpy_loops(x, w, (s, s), pb, pe, 'same_upper')
        b = ref_ops.groupconv_numpy(x, w, (s, s), pb, pe, 'same_upper')
        assert np.array_equal(a, b)


def test_mnist_known_answer():
    """README.md:69-72 / integrity_test.py:57: argsort == [2 0 1 7 8 6 3 4 5 9] on mnist2.png."""
    g = np.load(os.path.join(GOLDEN, 'mnist_e2e.npz'))
    names = json.loads(str(g['node_names']))
    from conftest import REPO
    for kt in ('numpy', 'special'):
        exe = ref_engine.load(os.path.join(REPO, 'models', 'mnist.xml'), kt)
        res = exe.infer({'conv2d_input': g['input']})
        prob = next(iter(res.values()))
        assert list(np.argsort(prob[0])[::-1]) == [2, 0, 1, 7, 8, 6, 3, 4, 5, 9]
        ok, msg = close(prob, g['final_' + kt], rtol=1e-5, atol=1e-7)
        assert ok, msg
        assert abs(float(prob[0, 2]) - 9.9999917e-01) < 1e-6          # README.md:69-71
        if kt == 'numpy':
            outs = exe.node_outputs()
            for i, nm in enumerate(names):
                ok, msg = close(outs[nm], g['node_{}'.format(i)], rtol=1e-5, atol=1e-6)
                assert ok, (nm, msg)
    exe = ref_engine.load(os.path.join(REPO, 'models', 'mnist.xml'), 'special', faithful_const=True)
    prob7 = next(iter(exe.infer({'conv2d_input': g['input7']}).values()))
    ok, msg = close(prob7, g['final7_special'], rtol=1e-5, atol=1e-7)
    assert ok, msg


def test_conv_pickle_known_answer():
    """resources/node_args_6.pickle (SSD conv0, real f16 weights) replayed through the oracle."""
    g = np.load(os.path.join(GOLDEN, 'conv_kat.npz'))
    data = json.loads(str(g['node']))['data']
    y = ref_ops.convolution(data, g['x_f16'], g['w_f16'], 'special', np.float16)
    assert y.shape == (1, 32, 150, 150) and y.dtype == np.float16
    same = hashlib.sha256(y.tobytes()).hexdigest() == str(g['sha_special'])
    # identical BLAS build -> identical bits; a different host may differ by one f16 ulp
    assert same or np.max(np.abs(y.astype(np.float32) - g['y_f16_special'].astype(np.float32))) <= 0.0157
    for kt in ('special', 'numpy'):
        y32 = ref_ops.convolution(data, g['x_f32_crop'], g['w_f32'], kt, np.float32)
        ok, msg = close(y32, g['y_f32_crop_' + kt], rtol=1e-5, atol=1e-6)
        assert ok, msg


E2E = os.path.join(GOLDEN, 'models_e2e.npz')


@pytest.mark.skipif(not os.path.isfile(E2E), reason='models_e2e.npz not generated')
@pytest.mark.parametrize('model,kt', [('mnist_bn', 'special'), ('mnist_bn', 'numpy'), ('googlenet-v1', 'special'),
                                      ('ssd_mobilenet_v1_coco', 'special')])
def test_models_end_to_end(model, kt, model_dir):
    from tools.synth_bin import synth_input
    g = np.load(E2E)
    x = synth_input(model, batch=2, seed=1)
    exe = ref_engine.load(os.path.join(model_dir, model + '.xml'), kt)
    name = exe.net.inputs[0]['name']
    for img in range(2):
        key = '{}|{}|{}'.format(model, kt, img)
        if key + '|final' not in g:
            continue
        res = next(iter(exe.infer({name: x[img:img + 1]}).values()))
        want = g[key + '|final']
        if model.startswith('ssd'):
            # records: [rank, class, conf, box]; class ids and order must be identical
            assert np.array_equal(res[..., 0:2], want[..., 0:2])
            ok, msg = close(res, want, rtol=1e-4, atol=1e-5)
        else:
            ok, msg = close(res, want, rtol=1e-4, atol=1e-7)
            assert np.argmax(res) == np.argmax(want)
        assert ok, (key, msg)
        if img == 0:
            names = json.loads(str(g[key + '|names']))
            outs = exe.node_outputs()
            samples = g[key + '|samples']
            absmax = g[key + '|absmax']
            for j, nm in enumerate(names):
                arr = np.asarray(outs[nm]).ravel()
                if arr.size == 0:
                    continue
                idx = np.arange(arr.size) if arr.size <= 64 else np.linspace(0, arr.size - 1, 64).astype(np.int64)
                got = np.resize(arr[idx].astype(np.float64), 64)
                tol = 1e-5 + 2e-5 * absmax[j]
                assert np.all(np.abs(got - samples[j]) <= tol + 1e-4 * np.abs(samples[j])), (model, nm)
