import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without CUDA skips the gpu tests instead of failing on the first one.  On a GPU host a
    missing libtmjx.so is NOT a skip: the product path must fail loudly there (the tests then error at `_lib.load()`)."""
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (no GPU on this host); run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def walker():
    from track_mjx_b200.walker import Rodent

    return Rodent(torque_actuators=True, rescale_factor=0.9)


@pytest.fixture(scope="session")
def clips2(walker):
    from track_mjx_b200 import clips

    return clips.make_synthetic_clips(walker.sections, 2)


@pytest.fixture(scope="session")
def task_cfg(walker):
    from track_mjx_b200 import config

    args = {k: v for k, v in config.DEFAULT_ENV_ARGS.items() if k != "reset_noise_scale"}
    return config.make_task_config(walker, config.RewardConfig(), **args)
