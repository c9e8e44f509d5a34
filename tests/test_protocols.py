"""Directory / protocol writers (generate_test_utterances.py:141-259) on a miniature corpus (CPU only)."""
import os
import random

import numpy as np
from scipy.io import wavfile

from spoofsv_b200 import protocols as P


def _mini_corpus(tmp_path, speakers=("p225", "p226", "p227"), n_real=5, n_syn=2):
    rng = np.random.default_rng(0)
    data = tmp_path / "data"
    syn = tmp_path / "spoof_data"
    for s in speakers:
        (data / "wav22" / s).mkdir(parents=True)
        (data / "txt" / s).mkdir(parents=True)
        for k in range(n_real):
            wavfile.write(data / "wav22" / s / f"{s}_{k + 1:03d}.wav", 22050, (rng.standard_normal(400) * 3000).astype(np.int16))
            (data / "txt" / s / f"{s}_{k + 1:03d}.txt").write_text(f"text of {s} number {k + 1}\nsecond line\n")
        (syn / ("s" + s[1:])).mkdir(parents=True)
        for k in range(n_syn):
            wavfile.write(syn / ("s" + s[1:]) / f"s{s[1:]}_{k + 1:03d}.wav", 22050, (0.1 * rng.standard_normal(2205)).astype(np.float32))
    return data, syn


def test_ivector_and_ge2e_layout(tmp_path):
    data, syn = _mini_corpus(tmp_path)
    sentences = ["the first sentence", "the second sentence"]
    out = tmp_path / "test" / "tag"
    counts = P.write_ivector_layout(str(data), str(syn), str(out), sentences, train_spk_num=2, enroll_utt_num=1,
                                    eval_utt_num=2, rng=random.Random(3))
    assert counts == {"train": 10, "test_real": 3, "test_spoof": 2}
    iv = out / "ivector_data"
    # training speakers: every real utterance, renumbered <spk>W<nnn>; the first one is also the dev set
    for spk in ("225", "226"):
        assert sorted(os.listdir(iv / "wav" / "train" / spk)) == [f"{spk}W{k:03d}.wav" for k in range(1, 6)]
    assert sorted(os.listdir(iv / "wav" / "dev")) == ["225"]
    assert sorted(os.listdir(iv / "wav" / "dev" / "225")) == sorted(os.listdir(iv / "wav" / "train" / "225"))
    # test speaker: enroll + eval real utterances, then the spoofed ones numbered after them
    assert sorted(os.listdir(iv / "wav" / "test" / "227")) == [f"227W{k:03d}.wav" for k in range(1, 6)]
    assert sorted(os.listdir(iv / "test_nospoof" / "227")) == [f"227W{k:03d}.wav" for k in range(1, 4)]
    sr, spoof = wavfile.read(iv / "wav" / "test" / "227" / "227W004.wav")
    sr0, src = wavfile.read(syn / "s227" / "s227_001.wav")
    assert sr == sr0 and np.array_equal(spoof, src)
    # transcripts: "<id>    <first line of the txt>" for real ones (both files), the TTS sentences for spoofed ones
    lines = (iv / "transcript" / "VCTK-transcript.txt").read_text().splitlines()
    ns = (iv / "VCTK-transcript_nospoof.txt").read_text().splitlines()
    assert len(lines) == 15 and len(ns) == 13 and lines[:13] == ns
    assert lines[13] == "227W004    the first sentence" and lines[14] == "227W005    the second sentence"
    by_id = dict(l.split("    ", 1) for l in ns)
    # the shuffled copy keeps wav and text together
    for k in range(1, 6):
        sr, w = wavfile.read(iv / "wav" / "train" / "226" / f"226W{k:03d}.wav")
        text = by_id[f"226W{k:03d}"]
        n = int(text.rsplit(" ", 1)[1])
        sr0, w0 = wavfile.read(data / "wav22" / "p226" / f"p226_{n:03d}.wav")
        assert np.array_equal(w, w0) and text == f"text of p226 number {n}"
    # GE2E: one symlink per speaker of train and test
    assert P.link_ge2e(str(out)) == 3
    ge = out / "ge2e_data"
    assert sorted(os.listdir(ge)) == ["225", "226", "227"] and all((ge / d).is_symlink() for d in os.listdir(ge))
    assert sorted(os.listdir(ge / "227")) == sorted(os.listdir(iv / "wav" / "test" / "227"))
    assert P.link_ge2e(str(out)) == 3          # idempotent


def test_antispoof_protocol(tmp_path):
    data, syn = _mini_corpus(tmp_path)
    anti = tmp_path / "antispoof"
    (anti / "ASVspoof2019_LA_cm_protocols").mkdir(parents=True)
    (anti / "ASVspoof2019_LA_dev" / "flac").mkdir(parents=True)
    dev_lines = []
    for k in range(4):
        (anti / "ASVspoof2019_LA_dev" / "flac" / f"LA_D_9{k}.flac").write_bytes(b"fLaC" + bytes([k]))
        dev_lines.append(f"LA_00{k} LA_D_9{k} - - bonafide")
    dev_lines.append("LA_0099 LA_D_77 - A01 spoof")
    (anti / "ASVspoof2019_LA_cm_protocols" / "ASVspoof2019.LA.cm.dev.trl.txt").write_text("\n".join(dev_lines) + "\n")
    counts = P.write_antispoof_set(str(anti), str(syn), "tag", bonafide_num=3)
    assert counts == {"bonafide": 3, "spoof": 6}
    proto = (anti / "ASVspoof2019_LA_cm_protocols" / "customized_data_tag.txt").read_text().splitlines()
    assert proto[:3] == [f"LA_00{k} LA_D_{k + 1:07d} - - bonafide" for k in range(3)]
    assert len(proto) == 9 and all(l.endswith("- - spoof") for l in proto[3:])
    assert {l.split()[0] for l in proto[3:]} == {"s225", "s226", "s227"}
    assert (anti / "tag" / "flac" / "LA_D_0000002.flac").read_bytes() == b"fLaC\x01"
    sr, w = wavfile.read(anti / "tag" / "flac" / (proto[3].split()[1] + ".wav"))
    assert sr == 16000 and w.dtype == np.int16 and len(w) == 1600          # 0.1 s at 22.05 kHz -> 16 kHz
    import pytest
    with pytest.raises(ValueError, match="not a bona-fide"):
        P.write_antispoof_set(str(anti), str(syn), "tag2", bonafide_num=5)


def test_resample_keeps_a_tone():
    t = np.arange(22050) / 22050.0
    y = np.sin(2 * np.pi * 440.0 * t).astype(np.float32)
    z = P.resample_to_16k(y, 22050)
    assert len(z) == 16000
    ref = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000.0)
    assert np.abs(z[200:-200] - ref[200:-200]).max() < 1e-3


def test_layouts_match_the_reference_code(tmp_path, monkeypatch):
    """The writers against a tree produced by the REFERENCE'S OWN lines (generate_test_utterances.py:142-259, executed
    unmodified by oracle/protocols_fixture.py in the build container): every path, symlink target, copied wav (SHA-1),
    both transcripts and the anti-spoofing protocol."""
    import hashlib
    import json
    from pathlib import Path
    from oracle import protocols_fixture as PF

    gold = json.loads((Path(__file__).parent / "golden" / "protocols_ref.json").read_text())
    cfg = PF.build_corpus(tmp_path)
    # the fixture run listed directories in sorted order (the reference shuffles the raw file-system order)
    real_listdir = os.listdir
    monkeypatch.setattr(P.os, "listdir", lambda p=".": sorted(real_listdir(p)))
    out = Path(cfg["SRC_ROOT_DIR"]) / "test" / PF.TAG
    syn = out / "spoof_data"
    P.write_ivector_layout(cfg["DATA_ROOT_DIR"], str(syn), str(out), PF.SENTENCES, PF.TRAIN_SPK, PF.ENROLL, PF.EVAL,
                           rng=random.Random(PF.SEED))
    P.link_ge2e(str(out))
    P.write_antispoof_set(cfg["ANTISPOOF_DIR"], str(syn), PF.TAG, bonafide_num=PF.BONAFIDE,
                          writer=lambda path, samples, sr: Path(path).write_bytes(b"resampled"), ext=".flac")
    monkeypatch.undo()
    snap = PF.snapshot(tmp_path, cfg)
    bona = {k: v for k, v in snap["files"].items() if "/flac/" in k and v != "resampled"}
    rest = {k: v for k, v in snap["files"].items() if k not in bona}
    assert len(bona) == gold["bonafide_count"]
    assert hashlib.sha1(json.dumps(sorted(bona.items())).encode()).hexdigest() == gold["bonafide_digest"]
    assert rest == gold["files"]
    assert snap["links"] == gold["links"]
    assert snap["texts"] == gold["texts"]
