"""TEST INFRASTRUCTURE ONLY: CPU (numpy) restatement of the feature front-end of the reference.

data/dataset.py:94-123 turns a wav into the training / teacher-forcing inputs of Text2Mel and SSRN with
librosa 0.7 (requirements.txt:4; not installed here and absent from /root/reference, so its published
algorithms are restated):

    speech, sr = librosa.core.load(path, sr=None, mono=True)                    (:94)   float32 samples
    speech, _ = librosa.effects.trim(speech, 22)                                 (:95)   framed RMS, 22 dB below the peak
    speech = append(speech[0], speech[1:] - PREEMPH * speech[:-1])               (:96)
    lin = abs(librosa.stft(speech, n_fft, hop_length))                           (:97)   hann, centred, reflect padding
    mel = dot(librosa.filters.mel(sr, n_fft, n_mels), lin)                       (:98-99) Slaney scale, area-normalised
    LOG_FEATURE: 20 log10(max(1e-5, .)) -> clip((. - REF_DB + MAX_DB) / MAX_DB, 1e-8, 1)     (:101-105)
    else:        (. / max(.)) ** NORM_POWER.ANALYSIS                             (:107-112)
    mel frames 0, 4, 8, ... (REDUCTION), lin cut to 4 * (frames // 4) columns     (:115-118)

Parity is "unpinned" against librosa itself; tests/test_features.py pins the filter bank against
transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") (an independent implementation written
to match librosa), the STFT against scipy.signal.stft, and the trim rule against a direct restatement.
"""
from __future__ import annotations

import numpy as np

from . import vocoder_oracle as V


def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): linear below 1 kHz (200/3 Hz per mel), log above (27 mels per factor 6.4)."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels) with its 0.7 defaults (fmin 0, fmax sr/2, htk False, norm 1):
    triangles on the Slaney mel scale, each scaled by 2 / (its bandwidth in Hz).  (n_mels, 1 + n_fft/2) float32."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def preemphasis(y: np.ndarray, coeff: float) -> np.ndarray:
    """data/dataset.py:96."""
    return np.append(y[0], y[1:] - coeff * y[:-1])


def features(wave: np.ndarray, sr: int, cfg: dict):
    """wav samples -> (reduced mel (n_mels, T // R), normalised linear spectrogram (1 + n_fft/2, R * (T // R)))."""
    n_fft, hop = cfg["STFT"]["FFT_LENGTH"], cfg["STFT"]["HOP_LENGTH"]
    n_mels, red = cfg["COARSE_MELSPEC"]["FREQ_BINS"], cfg["COARSE_MELSPEC"]["REDUCTION"]
    y = np.asarray(wave, dtype=np.float32)
    s, e = V.trim_bounds(y.astype(np.float64), top_db=22.0)
    y = y[s:e]
    y = preemphasis(y, np.float32(cfg["PREEMPH"])).astype(np.float32)
    lin = np.abs(V.stft(y.astype(np.float64), n_fft, hop, n_fft)).astype(np.float32)
    mel = mel_filterbank(sr, n_fft, n_mels) @ lin
    if cfg.get("LOG_FEATURE", False):
        mel = 20.0 * np.log10(np.maximum(1e-5, mel))
        lin = 20.0 * np.log10(np.maximum(1e-5, lin))
        mel_n = np.clip((mel - cfg["REF_DB"] + cfg["MAX_DB"]) / cfg["MAX_DB"], 1e-8, 1)
        lin_n = np.clip((lin - cfg["REF_DB"] + cfg["MAX_DB"]) / cfg["MAX_DB"], 1e-8, 1)
    else:
        p = cfg["NORM_POWER"]["ANALYSIS"]
        lin_n = (lin / np.max(lin)) ** p
        mel_n = (mel / np.max(mel)) ** p
    t4 = mel.shape[1] // red
    return mel_n[:, [red * k for k in range(t4)]].astype(np.float32), lin_n[:, :red * t4].astype(np.float32)
