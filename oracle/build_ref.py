"""Recipe for oracle/_ref: the UNMODIFIED reference modules, staged for the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/ttsmodel_oracle.py).

The reference (MingruiYuan/SpoofSV) is pure Python: "compiling" it means staging the one module the hot path
lives in, `models/TTSModel.py` (melSyn :234-300, SSRN :319-362), byte for byte from where it lies under
/root/reference into `oracle/_ref/models/` -- a git-ignored, NOT gpurun-ignored directory, so the file travels to
the GPU box with the snapshot like a built `.so`, while no reference source enters the repository history.

    python oracle/build_ref.py            # in the build container (needs /root/reference)

`bench.py --impl reference` and `bench.py`'s `extra.torch_eager_gpu` leg import it through `load()` below when it is
present (`cpu_baseline.kind = "reference"`), and fall back to the oracle port otherwise (`kind = "port"`).
The reference drivers (generate_test_utterances.py, synthesize.py) import librosa / soundfile / matplotlib, which
this image does not have; their AR loop (generate_test_utterances.py:105-116) is restated in `ar_loop()` below and
drives the unmodified module through its public forward() protocol.
"""
from __future__ import annotations

import hashlib
import importlib.util
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
DST = HERE / "_ref"
FILES = ["models/TTSModel.py"]


def build(verbose: bool = True) -> bool:
    """Stage the reference module; returns False (and leaves oracle/_ref alone) when /root/reference is absent."""
    if not REF_SRC.exists():
        return False
    for rel in FILES:
        src, dst = REF_SRC / rel, DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
    digest = {rel: hashlib.sha256((DST / rel).read_bytes()).hexdigest() for rel in FILES}
    (DST / "SHA256").write_text("".join(f"{h}  {rel}\n" for rel, h in digest.items()))
    if verbose:
        for rel, h in digest.items():
            print(f"oracle/_ref/{rel}  sha256 {h[:16]}")
    return True


def available() -> bool:
    return all((DST / rel).exists() for rel in FILES)


def load():
    """Import the staged reference module (its `device` global binds at import: hide the GPUs first for a CPU run)."""
    if not available():
        raise FileNotFoundError("oracle/_ref is not staged (run oracle/build_ref.py in the build container)")
    name = "_spoofsv_reference_TTSModel"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, DST / "models" / "TTSModel.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def models(R, sd1, sd2, cfg):
    """Reference melSyn / SSRN with the given state_dicts, eval mode, on the module's own `device`."""
    m1 = R.melSyn(vocab_len=cfg["vocab_len"], condition=True, spkemb_dim=cfg["spkemb_dim"], textemb_dim=cfg["textemb_dim"],
                  freq_bins=cfg["freq_bins"], hidden_dim=cfg["hidden_dim"])
    m2 = R.SSRN(freq_bins=cfg["freq_bins"], output_bins=cfg["output_bins"], ssrn_dim=cfg["ssrn_dim"])
    m1.load_state_dict(sd1, strict=True)
    m2.load_state_dict(sd2, strict=True)
    return m1.to(R.device).eval(), m2.to(R.device).eval()


def ar_loop(R, m1, textid, spkemb, n_frames, freq_bins=80):
    """generate_test_utterances.py:105-116 (the re-encoding AR loop), frame count as a parameter."""
    import torch
    B = textid.shape[0]
    init = torch.zeros((B, freq_bins, 1), device=R.device)
    Y, A, pma, K, V = m1(melspec=init, textid=textid, spkemb=spkemb, pma=torch.zeros((B,), device=R.device).long())
    inputs = torch.cat((init, Y), dim=-1)
    for _ in range(n_frames - 1):
        Y, A, pma = m1(melspec=inputs, textid=None, spkemb=spkemb, K=K, V=V, A_last=A, pma=pma)
        inputs = torch.cat((inputs, Y[:, :, -1:]), dim=-1)
    return Y, A, pma


if __name__ == "__main__":
    ok = build()
    print("staged" if ok else "skipped: /root/reference not present")
