import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def renderer_fp32():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tgtc_style_b200 as T
    return T.NerfRenderer(device="cuda:0", mode="fp32")


@pytest.fixture(scope="session")
def renderer_bf16():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tgtc_style_b200 as T
    return T.NerfRenderer(device="cuda:0", mode="bf16")
