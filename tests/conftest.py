import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def tiny_setup():
    """Config, weights and synthetic frame of the tiny golden case (regenerated from seeds)."""
    import torch
    from oracle import pr_oracle as O
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    lr, hr = O.synthetic_frame(cfg, 1)
    return cfg, sd, lr, hr
