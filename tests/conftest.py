import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "few-shot-transformer-tts_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def full_params():
    """Full-size synthetic weights (83.5 M parameters), rebuilt from seed 0."""
    from oracle import tts_oracle as O
    cfg = O.ModelConfig()
    return cfg, O.synth_params(cfg, seed=0)


@pytest.fixture(scope="session")
def tiny_params():
    from oracle import tts_oracle as O
    cfg = O.ModelConfig.tiny()
    return cfg, O.synth_params(cfg, seed=15)
