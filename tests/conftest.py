import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'hostemu'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
    import refenv
    # a device kernel that never returns cannot be interrupted from Python: bound every GPU test with a watchdog
    # thread that ends the process (a hung box is a lost GPU visit)
    for it in items:
        if 'gpu' in it.keywords and it.get_closest_marker('timeout') is None:
            it.add_marker(pytest.mark.timeout(240, method='thread'))
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:                          # a plain `pytest` on a GPU-less machine skips the GPU tests instead of failing
        skip_gpu = pytest.mark.skip(reason='no CUDA device (run with -m gpu on the B200 box)')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip_gpu)
    if not refenv.available():
        skip = pytest.mark.skip(reason='/root/reference not present (GPU box)')
        for it in items:
            if 'reference' in it.keywords:
                it.add_marker(skip)
