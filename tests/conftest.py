import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def same(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Numeric bit-exactness: equal values (-0 == +0) and NaNs in the same places."""
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.shape != b.shape:
        return False
    if a.is_floating_point():
        na, nb = torch.isnan(a), torch.isnan(b)
        return bool((na == nb).all()) and bool((a[~na] == b[~nb]).all())
    return bool((a == b).all())


@pytest.fixture(scope='session')
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
