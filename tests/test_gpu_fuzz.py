"""GPU: a short slice of the randomised parity sweep (profiles/fuzz_parity.py) as a regression test: random shapes
down to 3x3 coarse levels, 1-4 views and scales, every pose format / mask mode / depth convention, the disparity
head and both arithmetic modes against the float64 oracle."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu


def test_randomised_parity_slice():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'fuzz_parity.py')
    spec = importlib.util.spec_from_file_location('fuzz_parity', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(seed=5, n_cases=24, verbose=False) == 0


def test_randomised_flow_parity_slice():
    """The same for the fused flow-and-depth step (profiles/fuzz_flow.py): every pixel's gradient against the float32
    oracle, shapes off the 32 x 32 tile."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'fuzz_flow.py')
    spec = importlib.util.spec_from_file_location('fuzz_flow', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(seed=4, n_cases=16, verbose=False) == 0
