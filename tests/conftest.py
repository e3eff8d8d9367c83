# coding: utf-8
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def fixtures_pcm():
    z = np.load(GOLD / "fixtures_pcm.npz")
    return [z[f"pcm{i}"] for i in range(10)], z["n_frames"]


@pytest.fixture(scope="session")
def ref_fbank():
    z = np.load(GOLD / "ref_fbank.npz")
    return [z[f"fbank{i}"] for i in range(10)]


@pytest.fixture(scope="session")
def ref_cmvn():
    return np.load(GOLD / "ref_cmvn.npz")


@pytest.fixture(scope="session")
def ref_specaugment():
    return np.load(GOLD / "ref_specaugment.npz")


@pytest.fixture(scope="session")
def ref_processor():
    return np.load(GOLD / "ref_processor.npz")


@pytest.fixture(scope="session")
def ref_tables():
    return np.load(GOLD / "ref_tables.npz")
