import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference C (oracle/_ref).  Built here from /root/reference; on the GPU box the
    prebuilt library travels with the snapshot.  Tests that need it skip if it is absent."""
    from oracle import Ref, have_ref, build
    build()
    if not have_ref():
        pytest.skip("oracle/_ref/libx266ref.so not available")
    return Ref()


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def vectors():
    return dict(np.load(os.path.join(GOLDEN, "ref_vectors.npz")))


@pytest.fixture(scope="session")
def x266():
    """The product binding; the CUDA library must be built and loadable."""
    import x266_b200
    x266_b200.lib()
    return x266_b200
