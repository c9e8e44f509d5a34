"""Seeded weight / input builders shared by oracle/make_golden.py, tests/ and bench.py.

TEST INFRASTRUCTURE ONLY (see oracle/ttsmodel_oracle.py).  Weights are built with the product's
parameter containers (spoofsv_b200.models.TTSModel), whose construction order reproduces the
reference's RNG stream (checked against the reference in make_golden.py), so the same seed gives
the same state_dict here, on the GPU box, and in the unmodified reference.
"""
from __future__ import annotations

import hashlib
from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"

CFG = dict(vocab_len=34, spkemb_dim=200, textemb_dim=128, freq_bins=80, hidden_dim=256,
           output_bins=513, ssrn_dim=256)          # reference config.json:7-15,19-24


def build_models(seed: int = 0, init: str = "default", ln_jitter: bool = False):
    """torch.manual_seed(seed); melSyn first, SSRN second (SURVEY.md 8d cfg 1).

    init='kaiming' applies train/ordinary.py:16-19 (kaiming_normal_ on every >1-D weight);
    ln_jitter randomises LayerNorm affine parameters so they are not the trivial ones/zeros."""
    from spoofsv_b200.models import SSRN, melSyn
    torch.manual_seed(seed)
    m1 = melSyn(vocab_len=CFG["vocab_len"], condition=True, spkemb_dim=CFG["spkemb_dim"],
                textemb_dim=CFG["textemb_dim"], freq_bins=CFG["freq_bins"], hidden_dim=CFG["hidden_dim"])
    m2 = SSRN(freq_bins=CFG["freq_bins"], output_bins=CFG["output_bins"], ssrn_dim=CFG["ssrn_dim"])
    if init == "kaiming":
        def _init(layer):
            if hasattr(layer, "weight") and layer.weight is not None and layer.weight.dim() > 1:
                torch.nn.init.kaiming_normal_(layer.weight, nonlinearity="relu")
        m1.apply(_init)
        m2.apply(_init)
    elif init != "default":
        raise ValueError(init)
    if ln_jitter:
        g = torch.Generator().manual_seed(seed + 1000)
        for m in (m1, m2):
            for mod in m.modules():
                if isinstance(mod, torch.nn.LayerNorm):
                    with torch.no_grad():
                        mod.weight.copy_(torch.empty_like(mod.weight).uniform_(0.5, 1.5, generator=g))
                        mod.bias.copy_(torch.empty_like(mod.bias).uniform_(-0.5, 0.5, generator=g))
    m1.eval()
    m2.eval()
    return m1, m2


def state_dicts(seed: int = 0, init: str = "default", ln_jitter: bool = False) -> Tuple[Dict, Dict]:
    m1, m2 = build_models(seed, init, ln_jitter)
    return ({k: v.detach().clone() for k, v in m1.state_dict().items()},
            {k: v.detach().clone() for k, v in m2.state_dict().items()})


def sd_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def highway_params(d: int, k: int, seed: int):
    """Seeded highwayConv parameters with non-trivial LayerNorm affine."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / np.sqrt(d * k)
    u = lambda *shape, lo, hi: torch.empty(*shape).uniform_(lo, hi, generator=g)
    return {
        "conv.weight": u(2 * d, d, k, lo=-bound, hi=bound), "conv.bias": u(2 * d, lo=-bound, hi=bound),
        "ln1.weight": u(d, lo=0.5, hi=1.5), "ln1.bias": u(d, lo=-0.5, hi=0.5),
        "ln2.weight": u(d, lo=0.5, hi=1.5), "ln2.bias": u(d, lo=-0.5, hi=0.5),
    }


def load_fixtures():
    """Vendored inputs of the reference: spk_emb/*.npy (108 x 200) and havard.txt (720 lines)."""
    z = np.load(GOLDEN / "spk_emb.npz")
    names = [str(s) for s in z["names"]]
    emb = z["emb"].astype(np.float32)
    lines = [ln.strip() for ln in (GOLDEN / "havard.txt").read_text().splitlines()]
    return names, emb, lines


def synthetic_text(B: int, N: int, seed: int = 0) -> torch.Tensor:
    """SURVEY.md 8d cfg 3: ids uniform in [2, 33], last real id 1 ('E'), zero padding; lengths vary."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros((B, 1, N), dtype=torch.int64)
    for b in range(B):
        n = N if b == 0 else int(torch.randint(max(2, N // 2), N + 1, (1,), generator=g))
        ids[b, 0, : n - 1] = torch.randint(2, 34, (n - 1,), generator=g)
        ids[b, 0, n - 1] = 1
    return ids
