"""The reference's dataset and batch collation (data/dataset.py) on top of the GPU feature front-end.

Same constructor arguments, same path-list files (`<DATA_ROOT_DIR>data_path/ordinary/{wav,txt}.path.<mode>`, the
`ubm-finetune` lists), same sample dictionaries (`data_0` reduced mel, `data_1` text ids or the linear
spectrogram, `data_2` speaker embedding, `data_3` linear spectrogram) and the same `<spec_dir><spk>/<utt>_mel.npy`
/ `_lin.npy` cache.  The spectrograms come from `spoofsv_b200.features` (CUDA), so items are produced in the main
process (use `num_workers=0`).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np
import torch
from torch.utils.data import Dataset

from . import features as FT
from . import text as TX


class dataset(Dataset):
    """data/dataset.py:9-133."""

    def __init__(self, cfg, mode="train", pattern="conditional", step="train_text2mel", stage=None, spec_dir=None,
                 NAME="VCTK-Corpus"):
        self.cfg, self.mode, self.step, self.spec_dir, self.name = cfg, mode, step, spec_dir, NAME
        self.root_dir, self.spkemb_dir, self.vocabulary = cfg["DATA_ROOT_DIR"], cfg["SPK_EMB_DIR"], cfg["VOCABULARY"]
        if pattern in ("universal", "conditional"):
            base, tag = "data_path/ordinary/", mode
        elif pattern == "ubm-finetune" and stage in ("ubm", "finetune"):
            base, tag = "data_path/ubm-finetune/", f"{stage}.{mode}"
        else:
            raise ValueError(f"unknown pattern / stage: {pattern!r} / {stage!r}")
        self.wavlist = self._read_list(self.root_dir + base + "wav.path." + tag)
        self.txtlist = self._read_list(self.root_dir + base + "txt.path." + tag)

    @staticmethod
    def _read_list(path: str) -> List[str]:
        with open(path, "r") as f:
            return [ln.strip("\n") for ln in f.readlines()]

    def __len__(self):
        assert len(self.wavlist) == len(self.txtlist)
        return len(self.wavlist)

    def _spectrograms(self, idx: int):
        wav = self.wavlist[idx]
        if self.spec_dir is not None:
            return FT.cache_features(wav, self.spec_dir, self.cfg)
        y, sr = FT.load_wav(wav)
        mel, lin = FT.wav_features(y, sr, self.cfg)
        return mel.cpu().numpy(), lin.cpu().numpy()

    def __getitem__(self, idx):
        spk_id = self.wavlist[idx][-12:-8]                      # '.../p225/p225_001.wav' -> 'p225'
        mel, lin = self._spectrograms(idx)
        if self.step == "train_ssrn":
            return {"data_0": torch.from_numpy(mel), "data_1": torch.from_numpy(lin)}
        spk_emb = torch.from_numpy(np.expand_dims(np.load(self.spkemb_dir + spk_id + ".npy").astype(np.float32), axis=1))
        with open(self.txtlist[idx], "r") as f:
            line = f.readlines()[0].strip()
        ids = torch.from_numpy(TX.text2id(line, self.vocabulary))
        sample = {"data_0": torch.from_numpy(mel), "data_1": ids, "data_2": spk_emb}
        if not (self.step == "train_text2mel" or self.mode == "validate"):
            sample["data_3"] = torch.from_numpy(lin)
        return sample


def _pad_last(t: torch.Tensor, n: int) -> torch.Tensor:
    d = n - t.shape[-1]
    return t if d <= 0 else torch.cat((t, torch.zeros(t.shape[:-1] + (d,), dtype=t.dtype)), dim=-1)


def _collate(batch: List[Dict[str, torch.Tensor]], keys) -> Dict[str, torch.Tensor]:
    out = {}
    for k in keys:
        if k == "data_2":
            out[k] = torch.stack([b[k] for b in batch], dim=0)
        else:
            n = max(b[k].shape[-1] for b in batch)
            out[k] = torch.stack([_pad_last(b[k], n) for b in batch], dim=0)
    return out


def collate_pad_2(batch):
    """data/dataset.py:184-207: zero-pad mel and linear spectrograms to the batch maxima."""
    return _collate(batch, ("data_0", "data_1"))


def collate_pad_3(batch):
    """data/dataset.py:209-232: zero-pad text ids (id 0 = 'P') and mel spectrograms."""
    return _collate(batch, ("data_0", "data_1", "data_2"))


def collate_pad_4(batch):
    """data/dataset.py:234-263."""
    return _collate(batch, ("data_0", "data_1", "data_2", "data_3"))
