import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are the parity tests proper: they need a CUDA device AND the built library.  Without either they
    are skipped (not failed), so a plain `pytest tests` on a CPU host reports the CPU results cleanly."""
    import torch
    lib = os.path.join(ROOT, 'neuron_gan_b200', 'libngan_b200.so')
    reason = None
    if not torch.cuda.is_available():
        reason = 'needs a CUDA device (run on the B200 box with -m gpu)'
    elif not os.path.isfile(lib):
        reason = f'{lib} is not built (python -c "import __graft_entry__ as g; g.build()")'
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if 'gpu' in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import torch
    path = os.path.join(ROOT, 'tests', 'golden', 'pggan_step_golden.pt')
    return torch.load(path, weights_only=False)
