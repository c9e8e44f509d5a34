import sys
import warnings
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

warnings.filterwarnings("ignore", category=UserWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def cuda_models():
    """Seed-0 default-init models on cuda:0 plus their CPU state_dicts."""
    import torch
    from oracle import weights as W
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m1, m2 = W.build_models(0)
    sd1 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    sd2 = {k: v.detach().clone() for k, v in m2.state_dict().items()}
    return m1.cuda(), m2.cuda(), sd1, sd2


@pytest.fixture(scope="session")
def cuda_models_k():
    """Seed-7 kaiming-init, LayerNorm-jittered models on cuda:0 plus their CPU state_dicts."""
    import torch
    from oracle import weights as W
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m1, m2 = W.build_models(7, init="kaiming", ln_jitter=True)
    sd1 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    sd2 = {k: v.detach().clone() for k, v in m2.state_dict().items()}
    return m1.cuda(), m2.cuda(), sd1, sd2
