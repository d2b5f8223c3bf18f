import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference mounted (authoring container only)")


@pytest.fixture(scope="session")
def state_dict_w1():
    from project_morpheus_b200 import weights

    return weights.random_state_dict(0, "w1")


@pytest.fixture(scope="session")
def state_dict_default():
    from project_morpheus_b200 import weights

    return weights.random_state_dict(0, "default")


@pytest.fixture(scope="session")
def oracle_w1(state_dict_w1):
    import torch
    from oracle import snac_ref

    torch.set_grad_enabled(False)
    return snac_ref.SNAC.from_state_dict(state_dict_w1).eval()


@pytest.fixture(scope="session")
def ensure_lib():
    from project_morpheus_b200 import build

    return build.build()
