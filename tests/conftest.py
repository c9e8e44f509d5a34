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


def _write_driver_cfg(tmp_path, n_lines=6, speakers=("p225", "p226", "p227")):
    import json
    import numpy as np
    from oracle import weights as W
    from spoofsv_b200 import text as T
    names, emb, lines = W.load_fixtures()
    (tmp_path / "spk_emb").mkdir()
    for s in speakers:
        np.save(tmp_path / "spk_emb" / f"{s}.npy", emb[names.index(s)])
    (tmp_path / "texts.txt").write_text("\n".join(lines[:n_lines]) + "\n")
    cfg = {"DATA_ROOT_DIR": str(tmp_path / "no_corpus") + "/", "SPK_EMB_DIR": str(tmp_path / "spk_emb") + "/",
           "SRC_ROOT_DIR": str(tmp_path) + "/", "SPK_EMB_DIM": 200, "HIDDEN_DIM": 256, "TEXT_EMB_DIM": 128, "SSRN_DIM": 256,
           "VOCABULARY": T.DEFAULT_VOCABULARY, "MAX_TEXT_LEN": 186, "MAX_FRAME_NUM": 9,
           "STFT": {"FFT_LENGTH": 1024, "HOP_LENGTH": 256}, "COARSE_MELSPEC": {"REDUCTION": 4, "FREQ_BINS": 80},
           "INFERENCE_TEXT2MEL_MODEL": "absent.tar.pth", "INFERENCE_SSRN_MODEL": "absent.tar.pth",
           "TTS_TEXTS": str(tmp_path / "texts.txt")}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(cfg))
    return p, cfg


@pytest.fixture
def write_driver_cfg():
    """Factory: a config.json + texts + speaker embeddings laid out like the reference's, in a temp directory."""
    return _write_driver_cfg
