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


def _load_cases(path):
    z = np.load(path, allow_pickle=False)
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


@pytest.fixture(scope='session')
def golden():
    return _load_cases(GOLDEN)


@pytest.fixture(scope='session')
def golden_consist():
    """tests/golden/consist_golden.npz: the left-right loss of train_depth_then_cam_lr.py:211-340 incl. the depth
    consistency term, from the reference's own functions (tests/golden/make_golden_consist.py)."""
    return _load_cases(os.path.join(ROOT, 'tests', 'golden', 'consist_golden.npz'))['lr_consist']


@pytest.fixture(scope='session')
def golden_flow():
    """tests/golden/flow_golden.npz: the flow-and-depth loss of train_optflow_combine.py:138-240 from the reference's
    own functions (tests/golden/make_golden_flow.py)."""
    return _load_cases(os.path.join(ROOT, 'tests', 'golden', 'flow_golden.npz'))['flow']


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative' of BASELINE.json's gradient tolerance."""
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
