"""CPU: the oracle restatement against the known-answer vectors generated from the unmodified
reference (oracle/make_golden.py).  Floating point: tolerance 2e-5 max-abs (the oracle reproduced
the reference bit-for-bit in the build container; the slack covers a different CPU's vector ISA)."""
import numpy as np
import pytest
import torch

from oracle import ttsmodel_oracle as O
from oracle import weights as W

TOL = 2e-5


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_fixtures_present(golden_dir):
    names, emb, lines = W.load_fixtures()
    assert len(names) == 108 and emb.shape == (108, 200) and emb.dtype == np.float32
    assert names[0] == "p225" and len(lines) == 720
    assert abs(float(np.linalg.norm(emb[0])) - 0.92) < 0.05


def test_text2id_matches_reference_lengths(golden_dir):
    _, _, lines = W.load_fixtures()
    lens = np.load(golden_dir / "havard_id_lens.npy")
    got = np.array([O.text2id(s).shape[-1] for s in lines])
    assert np.array_equal(got, lens)
    assert lens.min() == 27 and lens.max() == 58 and lens[:20].max() == 50
    first = O.text2id(lines[0])[0]
    assert first.tolist()[:10] == [22, 10, 7, 2, 4, 11, 20, 5, 10, 2] and first[-1] == 1 and len(first) == 43
    assert O.text2id('a"b')[0].tolist() == [3, 33, 4, 1]         # '"' maps onto the id of "'"
    assert O.text2id("A&!b")[0].tolist() == [3, 4, 1]            # characters outside the vocabulary are dropped


def test_pad_text_ids():
    t = O.pad_text_ids([np.array([[3, 4, 1]]), np.array([[5, 1]])])
    assert t.shape == (2, 1, 3) and t.dtype == torch.int64
    assert t[1, 0].tolist() == [5, 1, 0]


def test_seeded_weights_match_reference_digest(golden_dir):
    want = dict(l.split() for l in (golden_dir / "digests.txt").read_text().splitlines())
    sd1, sd2 = W.state_dicts(0)
    assert len(sd1) == 214 and len(sd2) == 76
    assert W.sd_digest(sd1) == want["t2m_seed0"] and W.sd_digest(sd2) == want["ssrn_seed0"]
    k1, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
    assert W.sd_digest(k1) == want["t2m_seed7_kaiming_jit"] and W.sd_digest(k2) == want["ssrn_seed7_kaiming_jit"]


def test_highway_cases(golden_dir):
    z = np.load(golden_dir / "highway_cases.npz")
    for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
        p = {"." + kk: v for kk, v in W.highway_params(d, k, 100 + i).items()}
        x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
        y = O.highway_conv(x, p, "", dil, bool(causal))
        assert float((y - _t(z[f"y{i}"])).abs().max()) <= TOL, (d, k, dil, causal)


def test_ssrn_golden(golden_dir):
    z = np.load(golden_dir / "ssrn_seed7.npz")
    _, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
    with torch.no_grad():
        lin = O.ssrn(_t(z["mel"]), k2)
    assert lin.shape == (2, 513, 148)
    assert float((lin - _t(z["lin"])).abs().max()) <= TOL


def test_small_batch_reference_loop_and_incremental(golden_dir):
    z = np.load(golden_dir / "small_seed7.npz")
    k1, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
    ids, spk = _t(z["textid"]), _t(z["spk"])
    T = z["Y"].shape[-1]
    with torch.no_grad():
        Y, A, traj, lin = O.synthesize(k1, k2, ids, spk, T)
        K, V = O.text_encoder(ids, k1)
        iY, iA, itraj = O.ar_loop_incremental(k1, ids, spk, T)
    assert float((K - _t(z["K"])).abs().max()) <= TOL and float((V - _t(z["V"])).abs().max()) <= TOL
    assert float((Y - _t(z["Y"])).abs().max()) <= TOL
    assert float((A - _t(z["A"])).abs().max()) <= TOL
    assert float((lin - _t(z["lin"])).abs().max()) <= TOL
    assert np.array_equal(traj.numpy(), z["traj"])
    # the O(T) recurrence is the same function (fp32 reordering noise only)
    assert float((iY - _t(z["Y"])).abs().max()) <= 1e-4 and float((iA - _t(z["A"])).abs().max()) <= 1e-4
    assert np.array_equal(itraj.numpy(), z["traj"])


def test_cfg1_trajectory_and_mel(golden_dir):
    """BASELINE config 1: seed 0, Harvard #1, p225, 217 frames."""
    z = np.load(golden_dir / "cfg1_seed0.npz")
    sd1, sd2 = W.state_dicts(0)
    ids, spk = _t(z["textid"]), _t(z["spk"])
    assert z["traj"][:15, 0].tolist() == [0, 0, 0, 0, 0, 2, 2, 4, 4, 5, 6, 7, 9, 10, 12] and z["traj"][-1, 0] == 42
    with torch.no_grad():
        iY, iA, itraj = O.ar_loop_incremental(sd1, ids, spk, 217)
        lin = O.ssrn(_t(z["Y"]), sd2)
    assert np.array_equal(itraj.numpy(), z["traj"])
    assert float((iY - _t(z["Y"])).abs().max()) <= 1e-4
    assert float((iA - _t(z["A"])).abs().max()) <= 1e-4
    assert float((lin[:, :, ::4] - _t(z["lin_t4"])).abs().max()) <= TOL
    assert float((lin[:, :, :16] - _t(z["lin_first"])).abs().max()) <= TOL


def test_reference_loop_prefix_of_cfg1(golden_dir):
    """The faithful O(T^2) loop on the first 24 frames (kept short: it re-encodes the prefix)."""
    z = np.load(golden_dir / "cfg1_seed0.npz")
    sd1, _ = W.state_dicts(0)
    with torch.no_grad():
        Y, A, traj = O.ar_loop_reference(sd1, _t(z["textid"]), _t(z["spk"]), 24)
    assert np.array_equal(traj.numpy(), z["traj"][:24])
    assert float((Y - _t(z["Y"])[:, :, :24]).abs().max()) <= TOL


def test_window_mask_edges():
    """pma at 0, in the middle and at the end of the text: weights outside [pma, pma+2] are exactly 0."""
    sd1, _ = W.state_dicts(0)
    ids = W.synthetic_text(3, 9, seed=1)
    spk = torch.full((3, 200, 1), 0.05)
    with torch.no_grad():
        K, V = O.text_encoder(ids, sd1)
        for pma in ([0, 4, 8], [7, 6, 1]):
            Y, A, mx = O.melsyn_eval_call(sd1, torch.zeros(3, 80, 2), None, spk, K=K, V=V,
                                          A_last=torch.zeros(3, 9, 1), pma=torch.tensor(pma))
            for b, p in enumerate(pma):
                col = A[b, :, -1]
                inside = torch.zeros(9, dtype=torch.bool)
                inside[p:min(p + 2, 8) + 1] = True
                assert torch.all(col[~inside] == 0) and abs(float(col.sum()) - 1) < 1e-6
                assert p <= int(mx[b]) <= min(p + 2, 8)


def test_oracle_train_forward_against_reference_golden(golden_dir):
    """Train branch of melSyn.forward (models/TTSModel.py:263-273): oracle == reference output, bit for bit."""
    z = np.load(golden_dir / "train_seed7.npz")
    sd1, _ = W.state_dicts(7, init="kaiming", ln_jitter=True)
    with torch.no_grad():
        Y, A = O.melsyn_train_forward(sd1, torch.from_numpy(z["mel"]), torch.from_numpy(z["textid"]), torch.from_numpy(z["spk"]))
    assert float((Y - torch.from_numpy(z["Y"])).abs().max()) <= 2e-6
    assert float((A - torch.from_numpy(z["A"])).abs().max()) <= 2e-6
