import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')
MODEL_CACHE = os.environ.get('B200OV_MODEL_CACHE', '/tmp/b200ov_models')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200 box)')


@pytest.fixture(scope='session')
def model_dir():
    """Directory with the four IR models (real mnist.bin, synthetic .bin for the others)."""
    from tools.synth_bin import ensure_model
    for m in ('mnist', 'mnist_bn', 'googlenet-v1', 'ssd_mobilenet_v1_coco'):
        ensure_model(m, MODEL_CACHE)
    return MODEL_CACHE


def close(a, b, rtol=1e-4, atol=1e-5):
    """The north-star FP32 tolerance: |a-b| <= atol + rtol*|b| elementwise."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False, 'shape {} vs {}'.format(a.shape, b.shape)
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    if bad.any():
        i = np.argmax(err - (atol + rtol * np.abs(b)))
        return False, '{} of {} outside tolerance; worst |d|={:.3e} at ref={:.6e}'.format(
            int(bad.sum()), a.size, float(err.ravel()[i]), float(b.ravel()[i]))
    return True, 'max |d| = {:.3e}'.format(float(err.max()) if err.size else 0.0)
