import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-gaussian-splatting_b200")
ORACLE = os.path.join(ROOT, "oracle")
for p in (PKG, ORACLE, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def gs(cuda):
    """Our extension + autograd boundary.  Fails (does not skip) when it is not built."""
    import gaussian   # noqa: F401  built in-tree by 3d-gaussian-splatting_b200/build.py
    import renderer
    return gaussian, renderer


@pytest.fixture(scope="session")
def ref(cuda):
    """The reference's own CUDA build (oracle/_ref); tests that need it skip if it is absent."""
    import ref_pipeline
    g, r = ref_pipeline.load_reference()
    if g is None:
        pytest.skip("oracle/_ref reference build not present")
    return g, r
