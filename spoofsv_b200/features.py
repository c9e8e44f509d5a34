"""Feature front-end of the reference on the GPU (next row of the scope table).

data/dataset.py:94-123 turns every wav into the Text2Mel / SSRN training inputs on one CPU thread with librosa:
trim (22 dB), pre-emphasis, |STFT| (n_fft 1024, hop 256), mel projection (80 Slaney filters), per-utterance
max-normalisation raised to NORM_POWER.ANALYSIS (or the LOG_FEATURE mapping), every 4th mel frame.  Here the
trim bounds come from a framed RMS on the device, the STFT is a batched cuFFT call (torch.stft -- a library FFT,
as librosa's is), and everything after it is the library's own kernel pair behind `ssv_spec_features`
(magnitude + mel projection + maxima in one pass over the spectrogram, normalise + reduce in a second).
The same caching layout as the reference (`<spec_dir>/<spk>/<utt>_mel.npy`, `_lin.npy`, :116-119) is kept.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

from . import _lib
from .vocoder import trim_bounds


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    log = 15.0 + 27.0 * np.log(np.maximum(f, 1e-30) / 1000.0) / np.log(6.4)
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), 200.0 * m / 3.0)


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) with the 0.7 defaults the reference relies on (data/dataset.py:98):
    fmin 0, fmax sr / 2, Slaney mel scale, every triangle scaled to unit area.  (n_mels, 1 + n_fft / 2) float32."""
    bins = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    lo, mid, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (bins[None, :] - lo) / (mid - lo)
    falling = (hi - bins[None, :]) / (hi - mid)
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    return (tri * (2.0 / (hi - lo))).astype(np.float32)


def load_wav(path: str) -> Tuple[np.ndarray, int]:
    """librosa.core.load(path, sr=None, mono=True) for PCM wav files: float32 samples in [-1, 1), native rate."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype == np.int16:
        y = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        y = x.astype(np.float32) / 2147483648.0
    elif x.dtype == np.uint8:
        y = (x.astype(np.float32) - 128.0) / 128.0
    else:
        y = x.astype(np.float32)
    if y.ndim == 2:
        y = y.mean(axis=1)
    return y, int(sr)


_FB_CACHE = {}


def _filterbank_on(device, sr: int, n_fft: int, n_mels: int) -> torch.Tensor:
    key = (str(device), sr, n_fft, n_mels)
    if key not in _FB_CACHE:
        _FB_CACHE[key] = torch.from_numpy(mel_filterbank(sr, n_fft, n_mels)).to(device).contiguous()
    return _FB_CACHE[key]


def wav_features(wave, sr: int, cfg: dict):
    """One utterance: samples (numpy or tensor) -> (reduced mel (n_mels, T // R), normalised linear spectrogram
    (1 + n_fft / 2, R * (T // R))) as CUDA tensors -- `reduced_mel_spec`, `lin_spec_norm` of data/dataset.py:94-118."""
    if not torch.cuda.is_available():
        raise RuntimeError("wav_features: no CUDA device (there is no CPU path)")
    n_fft, hop = int(cfg["STFT"]["FFT_LENGTH"]), int(cfg["STFT"]["HOP_LENGTH"])
    n_mels, red = int(cfg["COARSE_MELSPEC"]["FREQ_BINS"]), int(cfg["COARSE_MELSPEC"]["REDUCTION"])
    y = torch.as_tensor(np.asarray(wave) if not torch.is_tensor(wave) else wave, dtype=torch.float32).cuda().reshape(-1)
    if y.numel() < n_fft // 2 + 1:
        raise ValueError(f"wav_features: {y.numel()} samples are too few for reflect padding of {n_fft // 2}")
    starts, ends = trim_bounds(y[None, :], top_db=22.0)
    y = y[starts[0]:ends[0]]
    if y.numel() < n_fft // 2 + 1:
        raise ValueError("wav_features: nothing left after trimming")
    z = torch.cat([y[:1], y[1:] - float(cfg["PREEMPH"]) * y[:-1]])
    window = torch.hann_window(n_fft, periodic=True, device=z.device, dtype=torch.float32)
    D = torch.stft(z, n_fft, hop_length=hop, win_length=n_fft, window=window, center=True, pad_mode="reflect",
                   return_complex=True).contiguous()                 # (F, T) complex64
    F, T = D.shape
    T4 = T // red
    ri = torch.view_as_real(D)
    fb = _filterbank_on(z.device, sr, n_fft, n_mels)
    lin = torch.empty((F, red * T4), device=z.device, dtype=torch.float32)
    mel = torch.empty((n_mels, T4), device=z.device, dtype=torch.float32)
    ws = torch.empty(F * T + n_mels * T + 2, device=z.device, dtype=torch.float32)
    log_feature = bool(cfg.get("LOG_FEATURE", False))
    _lib.check(_lib.load().ssv_spec_features(
        ri.data_ptr(), F, T, fb.data_ptr(), n_mels, int(log_feature), float(cfg["NORM_POWER"]["ANALYSIS"]),
        float(cfg.get("REF_DB", 20)), float(cfg.get("MAX_DB", 100)), red, lin.data_ptr(), mel.data_ptr(), ws.data_ptr(),
        _lib.current_stream_ptr()))
    return mel, lin


def cache_features(wav_path: str, spec_dir: str, cfg: dict):
    """The spectrogram cache of data/dataset.py:84-92,115-119: `<spec_dir>/<spk>/<utt>_mel.npy` and `_lin.npy`;
    returns the two arrays (loaded from the cache when both files exist)."""
    stem = os.path.splitext(wav_path)[0]
    key = os.path.join(os.path.basename(os.path.dirname(stem)), os.path.basename(stem))      # 'p225/p225_001'
    mel_f, lin_f = os.path.join(spec_dir, key + "_mel.npy"), os.path.join(spec_dir, key + "_lin.npy")
    if os.path.exists(mel_f) and os.path.exists(lin_f):
        return np.load(mel_f), np.load(lin_f)
    y, sr = load_wav(wav_path)
    mel, lin = wav_features(y, sr, cfg)
    mel, lin = mel.cpu().numpy(), lin.cpu().numpy()
    os.makedirs(os.path.dirname(mel_f), exist_ok=True)
    # several ranks (torchrun) may compute the same utterance at once: write to a private temporary file and
    # rename it into place, so that a reader never sees a half-written array
    for dst, arr in ((mel_f, mel), (lin_f, lin)):
        tmp = "{}.{}.{}.tmp.npy".format(dst, os.getpid(), os.environ.get("RANK", "0"))
        np.save(tmp, arr)
        os.replace(tmp, dst)
    return mel, lin
