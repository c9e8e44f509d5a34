"""TEST INFRASTRUCTURE ONLY: CPU (numpy, float64) restatement of the waveform stage of the reference.

generate_test_utterances.py:130-139 and synthesize.py:134-147 turn each predicted linear spectrogram into a wav
with librosa 0.7 (requirements.txt:4, not installed here and absent from /root/reference, so its published
algorithms are restated): per-utterance max-normalise, raise to NORM_POWER.RECONSTRUCTION / ANALYSIS,
`librosa.core.griffinlim(n_iter=64, hop_length, win_length)` (the "fast Griffin-Lim" of Perraudin et al. with
momentum 0.99, hann window, centred frames with reflect padding, random initial phases), de-emphasis
`scipy.signal.lfilter([1], [1, -PREEMPH])`, `librosa.effects.trim(top_db=30)`, a 9 s cap and peak
normalisation to 0.75.  librosa draws the initial phases from numpy's global RNG, so bit parity with a reference
run is impossible by construction; here the phases are an explicit argument, which makes the restatement and the
GPU implementation comparable sample by sample.  Parity is "unpinned" against librosa itself (not installable in
this container); it is pinned against scipy.signal.stft/istft and lfilter in tests/test_vocoder.py.
"""
from __future__ import annotations

import numpy as np


def hann(win_length: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True): the periodic Hann window librosa uses."""
    n = np.arange(win_length)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def stft(y: np.ndarray, n_fft: int, hop: int, win_length: int) -> np.ndarray:
    """librosa.core.stft(center=True, pad_mode='reflect', window='hann') -> (1 + n_fft/2, frames) complex."""
    w = np.zeros(n_fft)
    lo = (n_fft - win_length) // 2
    w[lo:lo + win_length] = hann(win_length)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    return np.fft.rfft(yp[idx] * w[:, None], axis=0)


def istft(D: np.ndarray, hop: int, win_length: int) -> np.ndarray:
    """librosa.core.istft(center=True): overlap-add, divide by the window sum-square where it exceeds tiny,
    drop n_fft/2 samples at both ends."""
    n_fft = 2 * (D.shape[0] - 1)
    n_frames = D.shape[1]
    w = np.zeros(n_fft)
    lo = (n_fft - win_length) // 2
    w[lo:lo + win_length] = hann(win_length)
    frames = np.fft.irfft(D, n=n_fft, axis=0) * w[:, None]
    out_len = n_fft + hop * (n_frames - 1)
    y = np.zeros(out_len)
    wss = np.zeros(out_len)
    for i in range(n_frames):
        y[i * hop:i * hop + n_fft] += frames[:, i]
        wss[i * hop:i * hop + n_fft] += w * w
    ok = wss > np.finfo(np.float32).tiny
    y[ok] /= wss[ok]
    return y[n_fft // 2: out_len - n_fft // 2]


def griffinlim(S: np.ndarray, angles0: np.ndarray, n_iter: int = 64, hop: int = 256, win_length: int = 1024,
               momentum: float = 0.99) -> np.ndarray:
    """librosa.core.griffinlim of 0.7 with the initial phases passed in (librosa: exp(2j*pi*rand))."""
    n_fft = 2 * (S.shape[0] - 1)
    angles = angles0.astype(np.complex128).copy()
    rebuilt = 0.0
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles, hop, win_length)
        rebuilt = stft(inverse, n_fft, hop, win_length)
        angles = rebuilt - (momentum / (1 + momentum)) * tprev
        angles = angles / (np.abs(angles) + 1e-16)
    return istft(S * angles, hop, win_length)


def deemphasis(x: np.ndarray, coeff: float) -> np.ndarray:
    """scipy.signal.lfilter([1], [1, -coeff], x): y[n] = x[n] + coeff * y[n-1]."""
    y = np.empty_like(x, dtype=np.float64)
    acc = 0.0
    for n in range(len(x)):
        acc = x[n] + coeff * acc
        y[n] = acc
    return y


def trim_bounds(y: np.ndarray, top_db: float = 30.0, frame_length: int = 2048, hop_length: int = 512):
    """librosa.effects.trim: [start, end) of the region whose frame RMS is within top_db of the loudest frame."""
    yp = np.pad(y, frame_length // 2, mode="reflect")
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    rms = np.sqrt(np.mean(yp[idx] ** 2, axis=0))
    ref = rms.max()
    db = 20.0 * np.log10(np.maximum(1e-5, rms)) - 20.0 * np.log10(np.maximum(1e-5, ref))
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return 0, 0
    return int(nz[0] * hop_length), int(min(len(y), (nz[-1] + 1) * hop_length))


def postprocess(pred_lin: np.ndarray, angles0: np.ndarray, cfg: dict, n_iter: int = 64) -> np.ndarray:
    """One utterance of generate_test_utterances.py:126-139: (513, 4T) in (0, 1) -> samples.  LOG_FEATURE (:126-128):
    dB mapping instead of the max-normalisation, and the signal is written without peak scaling (:139)."""
    log_feature = bool(cfg.get("LOG_FEATURE", False))
    if log_feature:
        pred_lin = np.power(10, 0.05 * (pred_lin * cfg["MAX_DB"] - cfg["MAX_DB"] + cfg["REF_DB"]))
    else:
        pred_lin = pred_lin / pred_lin.max()
    spec = pred_lin ** (cfg["NORM_POWER"]["RECONSTRUCTION"] / cfg["NORM_POWER"]["ANALYSIS"])
    sig = griffinlim(spec, angles0, n_iter, cfg["STFT"]["HOP_LENGTH"], cfg["STFT"]["FFT_LENGTH"])
    sig = deemphasis(sig, cfg["PREEMPH"])
    a, b = trim_bounds(sig, 30.0)
    sig = sig[a:b][: 9 * cfg["SAMPLING_RATE"]]
    return sig if log_feature else sig / np.max(sig) * 0.75
