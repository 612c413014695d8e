import os
import sys

import numpy as np
import pytest

# several CUDA streams that wait for each other on the device (ranks emulated on one GPU) must not share a hardware queue
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session', params=['tiny', 'edge'])
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope='session')
def golden_tiny():
    return load_golden('tiny')


@pytest.fixture(scope='session')
def gowalla():
    return load_golden('gowalla')


def rel_err(a, b):
    """Norm-wise relative error with an rms floor (never pure element-wise rtol)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    rms = np.sqrt(np.mean(b * b)) if b.size else 0.0
    return float(np.max(np.abs(a - b) / (np.abs(b) + rms + 1e-30))) if b.size else 0.0
