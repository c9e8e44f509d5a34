"""Waveform stage (next row of the scope table): the CPU restatement against scipy, the GPU path against it."""
import numpy as np
import pytest
import scipy.signal as ss

from oracle import vocoder_oracle as V

CFG = {"NORM_POWER": {"ANALYSIS": 0.6, "RECONSTRUCTION": 1.3}, "STFT": {"FFT_LENGTH": 1024, "HOP_LENGTH": 256},
       "PREEMPH": 0.97, "SAMPLING_RATE": 22050, "LOG_FEATURE": False}


def _spec(rng, frames=40):
    """A plausible magnitude spectrogram: |STFT| of a few decaying partials plus noise, scaled into (0, 1]."""
    n = 256 * (frames - 1)
    t = np.arange(n) / 22050.0
    y = sum(np.sin(2 * np.pi * f * t) * np.exp(-3 * t) for f in (220.0, 440.0, 1330.0)) + 0.05 * rng.standard_normal(n)
    S = np.abs(V.stft(y, 1024, 256, 1024))
    return S / S.max()


def test_oracle_stft_istft_lfilter_against_scipy():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(5000)
    D = V.stft(y, 1024, 256, 1024)
    _, _, Z = ss.stft(np.pad(y, 512, mode="reflect"), window="hann", nperseg=1024, noverlap=768, nfft=1024, boundary=None,
                      padded=False)
    assert np.abs(D - Z * V.hann(1024).sum()).max() < 1e-10
    back = V.istft(D, 256, 1024)
    assert len(back) == 256 * (D.shape[1] - 1) and np.abs(back - y[:len(back)]).max() < 1e-12
    x = rng.standard_normal(3000)
    assert np.abs(V.deemphasis(x, 0.97) - ss.lfilter([1], [1, -0.97], x)).max() < 1e-12


def test_oracle_griffinlim_converges_and_trim():
    rng = np.random.default_rng(1)
    S = _spec(rng)
    a0 = np.exp(2j * np.pi * rng.random(S.shape))
    err = []
    for it in (1, 8, 32):
        y = V.griffinlim(S, a0, n_iter=it)
        err.append(np.linalg.norm(np.abs(V.stft(y, 1024, 256, 1024)) - S) / np.linalg.norm(S))
    assert err[2] < err[1] < err[0] and err[2] < 0.2           # spectral convergence improves with iterations
    sig = np.concatenate([np.zeros(6000), np.sin(np.arange(20000) * 0.05), 1e-4 * np.ones(7000)])
    a, b = V.trim_bounds(sig, 30.0)
    assert 4000 <= a <= 6144 and 26000 <= b <= 28200


@pytest.mark.gpu
def test_gpu_waveform_stage_matches_oracle():
    import torch
    from spoofsv_b200 import vocoder as G
    rng = np.random.default_rng(2)
    specs = np.stack([_spec(rng, 36), _spec(rng, 36)])
    a0 = np.exp(2j * np.pi * rng.random(specs.shape))
    # Griffin-Lim alone, same initial phases: sample-level agreement (fp32 FFTs vs float64 numpy, 16 iterations),
    # for the fused kernels and for the chain of library FFT calls
    for fused in (True, False):
        y = G.griffin_lim(torch.from_numpy(specs).float().cuda(), n_iter=16, angles0=torch.from_numpy(a0), fused=fused).cpu().numpy()
        for k in range(2):
            want = V.griffinlim(specs[k], a0[k], n_iter=16)
            assert y[k].shape == want.shape
            assert np.abs(y[k] - want).max() <= 2e-3 * np.abs(want).max()
    # zero iterations = one inverse STFT; an odd frame count leaves the last frame without a partner
    for T in (35, 36, 4):       # 4 frames: the shortest signal a single reflection of the 512-sample padding covers
        sp = np.stack([_spec(rng, 36)[:, :T]])
        ph = np.exp(2j * np.pi * rng.random(sp.shape))
        y0 = G.griffin_lim(torch.from_numpy(sp).float().cuda(), n_iter=0, angles0=torch.from_numpy(ph)).cpu().numpy()[0]
        want0 = V.istft(sp[0] * ph[0], 256, 1024)
        assert y0.shape == want0.shape and np.abs(y0 - want0).max() <= 1e-5 * max(np.abs(want0).max(), 1e-3)
        y3 = G.griffin_lim(torch.from_numpy(sp).float().cuda(), n_iter=3, angles0=torch.from_numpy(ph)).cpu().numpy()[0]
        want3 = V.griffinlim(sp[0], ph[0], n_iter=3)
        assert np.abs(y3 - want3).max() <= 1e-3 * np.abs(want3).max()
    # de-emphasis scan kernel: long rows, several blocks, in agreement with lfilter
    x = rng.standard_normal((3, 333568)).astype(np.float32)
    got = G.deemphasis(torch.from_numpy(x).cuda(), 0.97).cpu().numpy()
    want = ss.lfilter([1], [1, -0.97], x.astype(np.float64), axis=1)
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    x1 = rng.standard_normal((2, 777)).astype(np.float32)                  # fewer samples than threads
    assert np.abs(G.deemphasis(torch.from_numpy(x1).cuda(), 0.5).cpu().numpy() - ss.lfilter([1], [1, -0.5], x1, axis=1)).max() < 1e-5
    # the whole stage of generate_test_utterances.py:130-139
    lin = np.clip(specs ** (0.6 / 1.3), 1e-4, 1.0)                          # what SSRN would hand over, in (0, 1]
    waves = G.postprocess(torch.from_numpy(lin).float().cuda(), CFG, n_iter=16, angles0=torch.from_numpy(a0))
    for k in range(2):
        want = V.postprocess(lin[k], a0[k], CFG, n_iter=16)
        assert abs(len(waves[k]) - len(want)) <= 512                        # trim bounds move in 512-sample hops
        n = min(len(waves[k]), len(want))
        assert abs(float(np.max(waves[k])) - 0.75) < 1e-6
        if len(waves[k]) == len(want):
            assert np.abs(waves[k][:n] - want[:n]).max() <= 5e-3
    # LOG_FEATURE branch (generate_test_utterances.py:126-128): dB mapping, no max-normalisation, no peak scaling
    log_cfg = dict(CFG, LOG_FEATURE=True, MAX_DB=100, REF_DB=20)
    lin_db = np.clip((20 * np.log10(np.maximum(specs, 1e-5)) - 20 + 100) / 100, 1e-8, 1.0)
    waves = G.postprocess(torch.from_numpy(lin_db).float().cuda(), log_cfg, n_iter=16, angles0=torch.from_numpy(a0))
    for k in range(2):
        want = V.postprocess(lin_db[k], a0[k], log_cfg, n_iter=16)
        assert abs(len(waves[k]) - len(want)) <= 512
        if len(waves[k]) == len(want):
            assert np.abs(waves[k] - want).max() <= 5e-3 * np.abs(want).max()
