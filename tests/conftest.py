import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libpcramp_ref.so (the compiled reference)")


@pytest.fixture(scope="session")
def oracle():
    from tests.harness import OracleLib
    return OracleLib()


@pytest.fixture(scope="session")
def ref():
    from tests.harness import RefLib, REF_PATH
    if not os.path.exists(REF_PATH):
        pytest.skip("oracle/_ref/libpcramp_ref.so not built (needs /root/reference)")
    return RefLib()


@pytest.fixture(scope="session")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pcramp_b200 import PcrampGpu
    g = PcrampGpu(0)
    yield g
    g.close()
