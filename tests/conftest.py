import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'vsl_golden.npz')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Case(dict):
    """One golden case: arrays as torch tensors, strings / flag dicts decoded."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


@pytest.fixture(scope='session')
def golden():
    z = np.load(GOLDEN, allow_pickle=False)
    cases = {}
    for key in z.files:
        case, name = key.split('/', 1)
        a = z[key]
        if a.dtype.kind == 'U':
            v = str(a)
            if name == 'flags':
                v = ast.literal_eval(v)
        else:
            v = torch.from_numpy(np.array(a))
        cases.setdefault(case, Case())[name] = v
    return cases


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative' of BASELINE.json's gradient tolerance."""
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
