"""CPU: the C-ABI library builds, loads, and exports exactly what include/spoofsv_b200.h declares.
No compute entry point is called here (no GPU in the build container)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from spoofsv_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "spoofsv_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ssv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.ssv_version() == 100
    assert isinstance(lib.ssv_last_error(), bytes)
    assert lib.ssv_launch_count(1) >= 0


def test_argument_errors_are_reported_before_any_cuda_call():
    lib = _lib.load()
    out = ctypes.c_void_p()
    st = lib.ssv_ssrn_create(None, None, None, 0, 80, 513, 128, ctypes.byref(out))
    assert st == 1 and b"ssrn_dim" in lib.ssv_last_error()
    st = lib.ssv_highway_conv_fwd(None, None, None, None, None, None, None, 1, 256, 4, 3, 1, 0, None, 0, None)
    assert st == 1 and b"null" in lib.ssv_last_error()
    with pytest.raises(ValueError):
        _lib.check(st)


def test_only_sm100a_code_in_library():
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    r = subprocess.run([exe, "--list-elf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0
    archs = set(re.findall(r"sm_\d+a?", r.stdout))
    assert archs == {"sm_100a"}, archs


def test_modules_keep_reference_interface():
    import inspect
    from spoofsv_b200.models import SSRN, highwayConv, melSyn
    assert list(inspect.signature(melSyn.__init__).parameters)[1:] == [
        "vocab_len", "condition", "spkemb_dim", "textemb_dim", "freq_bins", "hidden_dim"]
    assert list(inspect.signature(melSyn.forward).parameters)[1:] == [
        "melspec", "textid", "spkemb", "K", "V", "A_last", "pma"]
    assert list(inspect.signature(SSRN.__init__).parameters)[1:] == ["freq_bins", "output_bins", "ssrn_dim"]
    assert list(inspect.signature(highwayConv.__init__).parameters)[1:] == [
        "dimension", "kernel_size", "dilation", "causal"]
    m1 = melSyn(vocab_len=34, condition=True, spkemb_dim=200)
    m2 = SSRN(freq_bins=80, output_bins=513, ssrn_dim=256)
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    assert len(sd1) == 214 and len(sd2) == 76
    assert tuple(sd1["text_encoder.textemb_layer.W.weight"].shape) == (128, 34)
    assert tuple(sd1["audio_decoder.conv5.weight"].shape) == (80, 256, 1)
    assert tuple(sd1["text_encoder.hc3.conv.weight"].shape) == (1024, 512, 1)
    assert tuple(sd2["ups1.deconv.weight"].shape) == (256, 256, 2)
    assert tuple(sd2["conv3.weight"].shape) == (513, 512, 1)
    # a checkpoint dict in the reference's format round-trips
    m1b = melSyn(vocab_len=34, condition=True, spkemb_dim=200)
    m1b.load_state_dict({"model_state_dict": sd1}["model_state_dict"], strict=True)


def test_no_cpu_fallback():
    from spoofsv_b200.models import SSRN, melSyn
    m2 = SSRN(freq_bins=80, output_bins=513, ssrn_dim=256).eval()
    with pytest.raises(_lib.SsvError, match="no CPU path"):
        m2(torch.zeros(1, 80, 4))
    m1 = melSyn(vocab_len=34, condition=True, spkemb_dim=200).eval()
    with pytest.raises(_lib.SsvError, match="no CPU path"):
        m1(melspec=torch.zeros(1, 80, 1), textid=torch.zeros(1, 1, 5).long(), spkemb=torch.zeros(1, 200, 1),
           pma=torch.zeros(1).long())
    m1.train()                       # the train branch (teacher-forced forward) is CUDA-only as well
    with pytest.raises(_lib.SsvError, match="no CPU path"):
        m1(torch.zeros(1, 80, 4), torch.zeros(1, 1, 5).long(), torch.zeros(1, 200, 1))
