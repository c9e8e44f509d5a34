"""Waveform stage of the reference drivers on the GPU (next row of the scope table).

generate_test_utterances.py:130-139 / synthesize.py:134-147 run, per utterance and on one CPU thread:
max-normalise, `** (RECONSTRUCTION / ANALYSIS)`, 64 iterations of librosa's fast Griffin-Lim (n_fft 1024, hop 256),
de-emphasis, trim (30 dB), 9 s cap, peak 0.75, wav.  Here the whole batch is processed at once: Griffin-Lim in
the library's fused kernels (`ssv_griffin_lim`: shared-memory FFTs, overlap-add, momentum phase update; a chain of
torch.stft / torch.istft calls remains as the path for other STFT shapes and as a cross-check), de-emphasis in the
library's own scan kernel (`ssv_deemphasis`), trim bounds from framed RMS.
librosa seeds the phases from numpy's global RNG, so a reference run is not reproducible sample by sample;
`angles0` makes this implementation comparable with the CPU restatement in oracle/vocoder_oracle.py.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _lib


def griffin_lim(S: torch.Tensor, n_iter: int = 64, hop: int = 256, win_length: int = 1024, momentum: float = 0.99,
                angles0: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                fused: Optional[bool] = None) -> torch.Tensor:
    """(B, 1 + n_fft/2, frames) magnitudes on the GPU -> (B, hop * (frames - 1)) samples.
    librosa.core.griffinlim (0.7): momentum 0.99, hann window, centred reflect-padded frames.

    The reference's STFT shape (n_fft = win_length = 1024, hop = 256) runs in the library's fused kernels
    (`ssv_griffin_lim`, csrc/griffinlim.cu: shared-memory FFTs on frame pairs, momentum update in the STFT epilogue);
    other shapes, or fused=False, take the chain of torch.stft / torch.istft calls (cuFFT)."""
    _lib.require_cuda(S, "griffin_lim")
    B, F, T = S.shape
    n_fft = 2 * (F - 1)
    if angles0 is None:
        ph = torch.rand((B, F, T), device=S.device, generator=generator) * (2.0 * np.pi)
        angles = torch.polar(torch.ones_like(ph), ph)
    else:
        angles = angles0.to(device=S.device, dtype=torch.complex64)
    S = S.to(torch.float32).contiguous()
    length = hop * (T - 1)
    can_fuse = n_fft == 1024 and hop == 256 and win_length == 1024 and T >= 4      # single reflection: 256 (T - 1) > 512
    if fused is None:
        fused = can_fuse
    if fused:
        if not can_fuse:
            raise ValueError("griffin_lim: the fused kernels are built for n_fft = win_length = 1024, hop = 256 only")
        lib = _lib.load()
        n_ws = int(lib.ssv_griffin_lim_workspace(B, T))
        ws = torch.empty(n_ws, device=S.device, dtype=torch.float32)
        y = torch.empty((B, length), device=S.device, dtype=torch.float32)
        a_ri = torch.view_as_real(angles.contiguous())
        _lib.check(lib.ssv_griffin_lim(S.data_ptr(), a_ri.data_ptr(), B, F, T, int(n_iter), hop, win_length,
                                       float(momentum), y.data_ptr(), ws.data_ptr(), n_ws, _lib.current_stream_ptr()))
        return y
    window = torch.hann_window(win_length, periodic=True, device=S.device, dtype=torch.float32)
    istft = lambda D: torch.istft(D, n_fft, hop_length=hop, win_length=win_length, window=window, center=True, length=length)
    rebuilt = None
    c = momentum / (1.0 + momentum)
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles)
        rebuilt = torch.stft(inverse, n_fft, hop_length=hop, win_length=win_length, window=window, center=True,
                             pad_mode="reflect", return_complex=True)
        angles = rebuilt if tprev is None else rebuilt - c * tprev
        angles = angles / (angles.abs() + 1e-16)
    return istft(S * angles)


def deemphasis(x: torch.Tensor, coeff: float) -> torch.Tensor:
    """scipy.signal.lfilter([1], [1, -coeff]) over the last axis of a (B, n) CUDA tensor (ssv_deemphasis kernel)."""
    _lib.require_cuda(x, "deemphasis")
    x = x.to(torch.float32).contiguous()
    y = torch.empty_like(x)
    B, n = x.shape
    _lib.check(_lib.load().ssv_deemphasis(x.data_ptr(), y.data_ptr(), B, n, float(coeff), _lib.current_stream_ptr()))
    return y


def trim_bounds(y: torch.Tensor, top_db: float = 30.0, frame_length: int = 2048, hop_length: int = 512):
    """librosa.effects.trim bounds for every row of (B, n): lists of start / end sample indices."""
    B, n = y.shape
    yp = torch.nn.functional.pad(y[:, None, :], (frame_length // 2, frame_length // 2), mode="reflect")[:, 0, :]
    fr = yp.unfold(1, frame_length, hop_length)                     # (B, frames, frame_length)
    rms = fr.pow(2).mean(-1).sqrt()
    db = 20.0 * torch.log10(rms.clamp_min(1e-5)) - 20.0 * torch.log10(rms.max(dim=1, keepdim=True).values.clamp_min(1e-5))
    keep = (db > -top_db).cpu().numpy()
    starts, ends = [], []
    for k in range(B):
        nz = np.flatnonzero(keep[k])
        if nz.size == 0:
            starts.append(0); ends.append(0)
        else:
            starts.append(int(nz[0]) * hop_length); ends.append(int(min(n, (int(nz[-1]) + 1) * hop_length)))
    return starts, ends


def postprocess(pred_lin: torch.Tensor, cfg: dict, n_iter: int = 64, angles0: Optional[torch.Tensor] = None,
                generator: Optional[torch.Generator] = None) -> List[np.ndarray]:
    """generate_test_utterances.py:126-139 / synthesize.py:134-147 for a whole batch: (B, 513, 4T) in (0, 1) on the GPU
    -> one float32 waveform per utterance (trimmed, at most 9 s).  LOG_FEATURE false: per-utterance max-normalised
    magnitudes, output peak 0.75.  LOG_FEATURE true (:126-128): the prediction is a normalised dB value,
    magnitude = 10 ** (0.05 * (p * MAX_DB - MAX_DB + REF_DB)), no max-normalisation and no output scaling."""
    _lib.require_cuda(pred_lin, "postprocess")
    x = pred_lin.to(torch.float32)
    log_feature = bool(cfg.get("LOG_FEATURE", False))
    power = cfg["NORM_POWER"]["RECONSTRUCTION"] / cfg["NORM_POWER"]["ANALYSIS"]
    if log_feature:
        spec = torch.pow(10.0, 0.05 * (x * cfg["MAX_DB"] - cfg["MAX_DB"] + cfg["REF_DB"])).pow(power)
    else:
        spec = (x / x.amax(dim=(1, 2), keepdim=True)).pow(power)
    sig = griffin_lim(spec, n_iter, cfg["STFT"]["HOP_LENGTH"], cfg["STFT"]["FFT_LENGTH"], angles0=angles0, generator=generator)
    sig = deemphasis(sig, cfg["PREEMPH"])
    starts, ends = trim_bounds(sig, 30.0)
    cap = 9 * cfg["SAMPLING_RATE"]
    ends = [min(e, s + cap) for s, e in zip(starts, ends)]
    # peak normalisation (time_signal / np.max(time_signal) * 0.75 over the trimmed, capped signal) on the device, one
    # copy of the batch through a pinned buffer, then per-utterance slices
    B, L = sig.shape
    idx = torch.arange(L, device=sig.device)[None, :]
    lo = torch.tensor(starts, device=sig.device)[:, None]
    hi = torch.tensor(ends, device=sig.device)[:, None]
    inside = (idx >= lo) & (idx < hi)
    if log_feature:
        scaled = sig                                   # written as is (generate_test_utterances.py:139)
    else:
        peak = torch.where(inside, sig, torch.full_like(sig, -float("inf"))).amax(dim=1, keepdim=True)
        scaled = sig * (0.75 / peak)
    host = _pinned_like(scaled)
    host.copy_(scaled, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = host.numpy()
    return [arr[k, starts[k]:ends[k]].copy() if ends[k] > starts[k] else np.zeros(0, np.float32) for k in range(B)]


_PINNED = {}


def _pinned_like(t: torch.Tensor) -> torch.Tensor:
    """A reusable pinned host buffer of t's shape (one per shape: the corpus driver calls with a fixed batch shape)."""
    key = (tuple(t.shape), t.dtype)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) > 4:
            _PINNED.clear()
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        _PINNED[key] = buf
    return buf


def write_wav(path, samples: np.ndarray, sampling_rate: int) -> None:
    """librosa.output.write_wav(norm=False) of 0.7 == scipy.io.wavfile.write of the float32 samples."""
    from scipy.io import wavfile
    wavfile.write(str(path), int(sampling_rate), np.asarray(samples, dtype=np.float32))
