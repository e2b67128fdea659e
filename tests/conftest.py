import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_next: bring-up tests of code that is not on the product path yet; they need a CUDA "
                                       "device AND MFVI_TEST_NEXT=1 (python -m pytest tests -m gpu_next)")


def pytest_collection_modifyitems(config, items):
    import torch
    have_gpu = torch.cuda.is_available()
    skip = pytest.mark.skip(reason="no CUDA device")
    skip_next = pytest.mark.skip(reason="bring-up test: needs a CUDA device and MFVI_TEST_NEXT=1")
    try:
        import pytest_timeout  # noqa: F401
        have_timeout = True
    except ImportError:
        have_timeout = False
    for item in items:
        # no test may hang the suite: the multi-process ones (trial fan-out, gloo ranks) get 5 minutes, everything else 15
        if have_timeout and item.get_closest_marker("timeout") is None:
            multi = any(k in item.name for k in ("trial", "two_rank", "gloo", "bo_loop"))
            item.add_marker(pytest.mark.timeout(300 if multi else 900))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip)
        if "gpu_next" in item.keywords and not (have_gpu and os.environ.get("MFVI_TEST_NEXT") == "1"):
            item.add_marker(skip_next)
