import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


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


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both tiers (the CPU tier checks that it loads and exports its symbols)."""
    from bogp_b200 import build
    build.build()


# ---------------------------------------------------------------- shared synthetic configurations
@pytest.fixture(scope="session")
def cfg1():
    """Config 1: 3 tonals, 21-element VLA, truth (1.07 km, 71.11 m), 100 x 50 grid."""
    from bogp_b200.modes import INV_TONALS_3, REC_Z_21, pekeris_modes, synthetic_covariance
    modes = pekeris_modes(INV_TONALS_3, REC_Z_21)
    K = synthetic_covariance(modes, 71.11, 1.07)
    return dict(modes=modes, K=K, r=np.linspace(0.1, 8.0, 100), z=np.linspace(1.0, 200.0, 50), truth=(1.07, 71.11))


@pytest.fixture(scope="session")
def cfg2_small():
    """Config 2 shape (13 tonals, 64 elements, truth (5 km, 60 m)) on a grid the oracle finishes in seconds."""
    from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    modes = pekeris_modes(SWELLEX_TONALS_13)
    K = synthetic_covariance(modes, 60.0, 5.0)
    return dict(modes=modes, K=K, r=np.linspace(0.1, 10.0, 160), z=np.linspace(1.0, 200.0, 24), truth=(5.0, 60.0))
