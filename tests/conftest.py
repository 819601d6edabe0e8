import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')
    # ICQB_EMULATE=1 (build container, no GPU): run `-m gpu` tests against the CPU emulation of
    # the library (tools/emu/run_emulated.py) -- a dry run of the tests' own code and of the host
    # layer before GPU time is spent on them; slow, so pick small cases with -k
    if os.environ.get('ICQB_EMULATE') == '1':
        sys.path.insert(0, os.path.join(ROOT, 'tools', 'emu'))
        import run_emulated
        run_emulated.install()


@pytest.fixture(scope='session')
def golden():
    import numpy

    class G:
        def __init__(self):
            self._cache = {}

        def __getitem__(self, name):
            if name not in self._cache:
                self._cache[name] = numpy.load(os.path.join(GOLDEN, name + '.npz'))
            return self._cache[name]

        def mesh(self, name):
            m = self['meshes']
            return m[name + '_xyz'], m[name + '_tri']
    return G()


@pytest.fixture(scope='session')
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope='session')
def gpu_lib():
    """The CUDA library; fails (not skips) when it is missing on a GPU box."""
    import torch
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    from icqsol_b200 import _lib
    return _lib.load()
