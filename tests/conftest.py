import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import nixoracle as no
    no.build(("port",))
    return no.load("port")


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import nixoracle as no
    if os.path.isdir("/root/reference"):
        no.build(("ref",))
    if not no.available("ref"):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return no.load("ref")


@pytest.fixture(scope="session")
def gpu_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from nix_b200 import core
    return core.load_library()  # raises (does not skip) when the CUDA extension is missing
