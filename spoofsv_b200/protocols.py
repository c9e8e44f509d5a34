"""Directory / protocol writers of the corpus driver (next row of the scope table).

generate_test_utterances.py:141-259 lays the synthesised wavs out for the (out-of-scope) consumers: a Kaldi
i-vector tree with transcripts (:141-218), a GE2E directory of symlinks (:220-227) and an ASVspoof-style
anti-spoofing set with a protocol file (:229-259).  The reference does this with `os.system('cp ...')` per file;
here it is plain Python file I/O with the same names, numbering and transcript lines:

    <root>/ivector_data/wav/train/<spk>/<spk>W<nnn>.wav        every real utterance of the first train_spk_num speakers
    <root>/ivector_data/wav/dev/<spk0>/...                      copy of the first training speaker
    <root>/ivector_data/wav/test/<spk>/<spk>W<nnn>.wav          enroll + eval real utterances, then the spoofed ones
    <root>/ivector_data/test_nospoof/<spk>/<spk>W<nnn>.wav      the real ones only
    <root>/ivector_data/transcript/VCTK-transcript.txt          "<utt id>    <text>" per line
    <root>/ivector_data/VCTK-transcript_nospoof.txt
    <root>/ge2e_data/<spk> -> ivector_data/wav/{train,test}/<spk>
    <ANTISPOOF_DIR>/<time>/flac/LA_D_<7 digits>.<ext> + ASVspoof2019_LA_cm_protocols/customized_data_<time>.txt

`<spk>` is the speaker directory name without its first character (p225 -> 225), as the reference writes it.
The reference shuffles each speaker's utterance list with the unseeded global `random`; here the generator is an
argument.  The anti-spoofing audio is resampled to 16 kHz (polyphase, scipy) and written by a caller-supplied
writer: the reference writes FLAC through `soundfile`, which this image does not have -- the default writer
stores 16-bit PCM wav under the same stem and the protocol lines are unchanged.
Parity: tests/test_protocols.py compares every path, link, copied file and text with a tree written by the reference's
own lines (tests/golden/protocols_ref.json, made by oracle/protocols_fixture.py).
"""
from __future__ import annotations

import os
import random
import shutil
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def _first_line(path: Path) -> str:
    with open(path, "r") as f:
        return f.readline().strip()


def _utt_id(spk: str, j: int) -> str:
    """'<spk>W<nnn>' with 1-based, zero-padded numbering (generate_test_utterances.py:168,193,212)."""
    return f"{spk}W{str(j + 1).zfill(3)}"


def write_ivector_layout(data_root: str, syn_root: str, out_root: str, sentences: Sequence[str], train_spk_num: int,
                         enroll_utt_num: int, eval_utt_num: int, rng: Optional[random.Random] = None) -> Dict[str, int]:
    """generate_test_utterances.py:141-218.  data_root holds wav22/<pXXX>/*.wav and txt/<pXXX>/*.txt, syn_root the
    synthesised s<XXX>/s<XXX>_<nnn>.wav; returns counts of what was written."""
    rng = rng or random.Random()
    data_root, syn_root, iv = Path(data_root), Path(syn_root), Path(out_root) / "ivector_data"
    real_root, txt_root = data_root / "wav22", data_root / "txt"
    real = sorted(os.listdir(real_root))
    syn = sorted(os.listdir(syn_root))
    if len(syn) < len(real):
        raise ValueError(f"{len(real)} real speakers but only {len(syn)} synthesised ones under {syn_root}")
    (iv / "transcript").mkdir(parents=True, exist_ok=True)
    counts = {"train": 0, "test_real": 0, "test_spoof": 0}
    with open(iv / "transcript" / "VCTK-transcript.txt", "w") as tr, open(iv / "VCTK-transcript_nospoof.txt", "w") as tr_ns:
        for i, spk_dir in enumerate(real):
            spk = spk_dir[1:]
            if spk != syn[i][1:]:
                raise ValueError(f"speaker order mismatch: real {spk_dir} vs synthesised {syn[i]}")
            utts = os.listdir(real_root / spk_dir)
            rng.shuffle(utts)
            if i < train_spk_num:
                dst = iv / "wav" / "train" / spk
                dst.mkdir(parents=True, exist_ok=True)
                for j, u in enumerate(utts):
                    shutil.copyfile(real_root / spk_dir / u, dst / (_utt_id(spk, j) + ".wav"))
                    line = f"{_utt_id(spk, j)}    {_first_line(txt_root / spk_dir / (u[:-4] + '.txt'))}\n"
                    tr.write(line)
                    tr_ns.write(line)
                    counts["train"] += 1
                if i == 0:
                    dev = iv / "wav" / "dev"
                    dev.mkdir(parents=True, exist_ok=True)
                    shutil.copytree(dst, dev / spk, dirs_exist_ok=True)
            else:
                dst, dst_ns = iv / "wav" / "test" / spk, iv / "test_nospoof" / spk
                dst.mkdir(parents=True, exist_ok=True)
                dst_ns.mkdir(parents=True, exist_ok=True)
                n_real = enroll_utt_num + eval_utt_num
                if len(utts) < n_real:
                    raise ValueError(f"speaker {spk_dir} has {len(utts)} utterances, {n_real} needed")
                for j in range(n_real):
                    u = utts[j]
                    shutil.copyfile(real_root / spk_dir / u, dst / (_utt_id(spk, j) + ".wav"))
                    shutil.copyfile(real_root / spk_dir / u, dst_ns / (_utt_id(spk, j) + ".wav"))
                    line = f"{_utt_id(spk, j)}    {_first_line(txt_root / spk_dir / (u[:-4] + '.txt'))}\n"
                    tr.write(line)
                    tr_ns.write(line)
                    counts["test_real"] += 1
                spoofed = sorted(os.listdir(syn_root / ("s" + spk)), key=lambda x: x[:-4])
                for j in range(eval_utt_num):          # numbered after the real ones: the Kaldi scripts rely on it
                    shutil.copyfile(syn_root / ("s" + spk) / spoofed[j], dst / (_utt_id(spk, j + n_real) + ".wav"))
                    tr.write(f"{_utt_id(spk, j + n_real)}    {sentences[j]}\n")
                    counts["test_spoof"] += 1
    return counts


def link_ge2e(out_root: str) -> int:
    """generate_test_utterances.py:220-227: ge2e_data/<spk> symlinks to every train and test speaker directory."""
    out_root = Path(out_root)
    ge2e = out_root / "ge2e_data"
    ge2e.mkdir(parents=True, exist_ok=True)
    n = 0
    for part in ("train", "test"):
        src = out_root / "ivector_data" / "wav" / part
        if not src.is_dir():
            continue
        for d in sorted(os.listdir(src)):
            link = ge2e / d
            if link.is_symlink() or link.exists():
                link.unlink()
            os.symlink(src / d, link)
            n += 1
    return n


def resample_to_16k(x: np.ndarray, sr: int) -> np.ndarray:
    """librosa.load(path, sr=16000) resamples with a windowed-sinc filter; here: scipy's polyphase resampler."""
    from math import gcd
    from scipy.signal import resample_poly
    if sr == 16000:
        return np.asarray(x, dtype=np.float32)
    g = gcd(16000, sr)
    return resample_poly(np.asarray(x, dtype=np.float64), 16000 // g, sr // g).astype(np.float32)


def _write_wav16(path: Path, samples: np.ndarray, sr: int) -> None:
    from scipy.io import wavfile
    wavfile.write(path, sr, np.clip(np.round(samples * 32767.0), -32768, 32767).astype(np.int16))


def write_antispoof_set(antispoof_dir: str, syn_root: str, current_time: str, bonafide_num: int = 10 * 108,
                        writer: Callable[[Path, np.ndarray, int], None] = _write_wav16, ext: str = ".wav") -> Dict[str, int]:
    """generate_test_utterances.py:229-259: the first bonafide_num bona-fide trials of the ASVspoof 2019 LA dev
    protocol are copied, every synthesised utterance is resampled to 16 kHz and appended as a spoof trial."""
    from .features import load_wav
    root = Path(antispoof_dir)
    out = root / current_time / "flac"
    out.mkdir(parents=True, exist_ok=True)
    proto_dir = root / "ASVspoof2019_LA_cm_protocols"
    with open(proto_dir / "ASVspoof2019.LA.cm.dev.trl.txt", "r") as f:
        dev = f.readlines()
    if len(dev) < bonafide_num:
        raise ValueError(f"the dev protocol holds {len(dev)} lines, {bonafide_num} bona-fide trials needed")
    index = 0
    with open(proto_dir / f"customized_data_{current_time}.txt", "w") as proto:
        for _ in range(bonafide_num):
            info = dev[index].strip().split()
            if info[-1] != "bonafide":
                raise ValueError(f"line {index + 1} of the dev protocol is not a bona-fide trial")
            name = f"LA_D_{str(index + 1).zfill(7)}"
            shutil.copyfile(root / "ASVspoof2019_LA_dev" / "flac" / (info[1] + ".flac"), out / (name + ".flac"))
            proto.write(f"{info[0]} {name} - - bonafide\n")
            index += 1
        n_spoof = 0
        for spk in os.listdir(syn_root):
            for utt in os.listdir(Path(syn_root) / spk):
                y, sr = load_wav(str(Path(syn_root) / spk / utt))
                name = f"LA_D_{str(index + 1).zfill(7)}"
                writer(out / (name + ext), resample_to_16k(y, sr), 16000)
                proto.write(f"{spk} {name} - - spoof\n")
                index += 1
                n_spoof += 1
    return {"bonafide": bonafide_num, "spoof": n_spoof}
