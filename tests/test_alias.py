"""The `pyopenvino` import alias (SURVEY.md 8(b) "plugin discovery"): scripts written for the reference resolve
`from pyopenvino.inference_engine import IECore`, top-level `common_def` and `op_plugins.<Type>` against this
package without edits.  Run in subprocesses from the repo root, like the reference's scripts are."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REPO


def _run(code_or_path, *args, is_path=False):
    if is_path:
        # the reference's scripts sit in the repo root, so python puts the root (their directory) on sys.path; ours sit in
        # tests/scripts/: run them unedited with sys.path[0] = repo root, exactly what a root-level script gets
        boot = ('import runpy, sys; sys.path.insert(0, {root!r}); sys.argv = sys.argv[1:]; '
                'runpy.run_path(sys.argv[0], run_name="__main__")').format(root=REPO)
        cmd = [sys.executable, '-c', boot, code_or_path] + list(args)
    else:
        cmd = [sys.executable, '-c', code_or_path] + list(args)
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    return subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=600)


def test_reference_import_lines_resolve():
    code = '\n'.join([
        'import sys',
        'from pyopenvino.inference_engine import IECore',          # test_pyopenvino.py:7
        'import pyopenvino_b200.inference_engine as impl',
        'assert IECore is impl.IECore',
        "sys.path.append('./pyopenvino')",                          # pyopenvino/inference_engine.py:17
        'import common_def',                                        # :18
        "assert common_def.type_convert_tbl['FP32'].__name__ == 'float32'",
        "assert common_def.string_to_tuple('1,2') == (1, 2)",
        'import op_plugins.Convolution as op',                      # test_node_sample.py:11
        'assert callable(op.compute) and callable(op.name)',
        'import pyopenvino.op_plugins.MaxPool as mp',               # what importlib does for the reference's Plugins
        'import pyopenvino_b200.op_plugins.MaxPool as mp2',
        'assert mp is mp2',
        'ie = IECore()',
        "assert 'Convolution' in ie.plugins.plugins and hasattr(ie.plugins, 'MatMul')",
        "net = ie.read_network('models/mnist.xml', 'models/mnist.bin')",
        "assert net.outputs[0]['name'] and net.inputs[0]['data']['shape'] == (1, 1, 28, 28)",
        "exenet = ie.load_network(net, 'CPU', num_requests=1)",
        "exenet.kernel_type = 'numpy'",
        "print('ALIAS-OK')",
    ])
    r = _run(code)
    assert r.returncode == 0 and 'ALIAS-OK' in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_style_script_runs_unedited(tmp_path):
    """tests/scripts/ref_style_mnist.py only knows the reference's names; README.md:69-72 known answer."""
    g = np.load(os.path.join(GOLDEN, 'mnist_e2e.npz'))
    img = tmp_path / 'mnist2.npy'
    np.save(img, g['input'].reshape(28, 28).astype(np.uint8))
    assert np.array_equal(g['input'].reshape(28, 28).astype(np.uint8).astype(np.float32), g['input'].reshape(28, 28))
    r = _run(os.path.join('tests', 'scripts', 'ref_style_mnist.py'), str(img), is_path=True)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith('RESULT')][0]
    assert line.split()[1:] == ['2', '0', '1', '7', '8', '6', '3', '4', '5', '9']


@pytest.mark.gpu
def test_reference_style_node_script(tmp_path):
    """test_node_sample.py style: host arrays in, host array out through `op_plugins.MaxPool`, checked against the oracle."""
    from oracle import ref_ops
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 8, 9, 9)).astype(np.float32)
    node = {'name': 'pool', 'type': 'MaxPool', 'version': 'opset1',
            'data': {'kernel': '3,3', 'strides': '2,2', 'pads_begin': '0,0', 'pads_end': '0,0', 'rounding_type': 'ceil', 'auto_pad': 'explicit'},
            'input': {0: {'precision': 'FP32', 'dims': x.shape}}, 'output': {1: {'precision': 'FP32', 'dims': (1, 8, 4, 4)}}}
    pk = tmp_path / 'node_args.pickle'
    with open(pk, 'wb') as f:
        pickle.dump((node, {0: x}), f)
    out = tmp_path / 'out.npy'
    r = _run(os.path.join('tests', 'scripts', 'ref_style_node.py'), str(pk), str(out), is_path=True)
    assert r.returncode == 0, r.stdout + r.stderr
    want = ref_ops.maxpool(node['data'], x)
    assert np.array_equal(np.load(out), want)
