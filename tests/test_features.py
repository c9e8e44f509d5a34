"""Feature front-end (data/dataset.py:94-123): the oracle against independent implementations (CPU), the GPU path
against the oracle."""
import numpy as np
import pytest

from oracle import features_oracle as FO
from oracle import vocoder_oracle as VO

CFG = {"STFT": {"FFT_LENGTH": 1024, "HOP_LENGTH": 256}, "COARSE_MELSPEC": {"FREQ_BINS": 80, "REDUCTION": 4},
       "PREEMPH": 0.97, "NORM_POWER": {"ANALYSIS": 0.6, "RECONSTRUCTION": 1.3}, "LOG_FEATURE": False,
       "MAX_DB": 100, "REF_DB": 20, "SAMPLING_RATE": 22050}


from oracle.testsignals import utterance as _utterance


# ------------------------------------------------------------------------------------------- CPU: pin the oracle
def test_mel_filterbank_matches_independent_implementation():
    """librosa.filters.mel restated vs transformers' mel_filter_bank (written to reproduce librosa, Slaney scale
    and Slaney area normalisation)."""
    from transformers.audio_utils import mel_filter_bank
    for sr, n_fft, n_mels in [(22050, 1024, 80), (16000, 512, 40)]:
        ref = mel_filter_bank(1 + n_fft // 2, n_mels, 0.0, sr / 2.0, sr, norm="slaney", mel_scale="slaney").T
        got = FO.mel_filterbank(sr, n_fft, n_mels)
        assert got.shape == (n_mels, 1 + n_fft // 2) and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 1e-8
        # every triangle has unit area in Hz (Slaney normalisation), away from the band edges
        area = got.astype(np.float64).sum(axis=1) * (sr / n_fft)
        assert np.allclose(area[5:-1], 1.0, atol=0.06)       # sampled triangles: a few percent off for narrow bands


def test_product_filterbank_equals_oracle():
    from spoofsv_b200.features import mel_filterbank
    for sr, n_fft, n_mels in [(22050, 1024, 80), (16000, 512, 40), (22050, 2048, 128)]:
        assert np.abs(mel_filterbank(sr, n_fft, n_mels) - FO.mel_filterbank(sr, n_fft, n_mels)).max() <= 1e-9


def test_oracle_stft_magnitude_matches_scipy():
    from scipy import signal
    y = _utterance(1)[:20000].astype(np.float64)
    D = VO.stft(y, 1024, 256, 1024)
    # scipy: boundary/padding off, hann periodic, unscaled spectrum of the same reflect-padded signal
    yp = np.pad(y, 512, mode="reflect")
    _, _, Z = signal.stft(yp, window=signal.get_window("hann", 1024, fftbins=True), nperseg=1024, noverlap=768,
                          boundary=None, padded=False, return_onesided=True)
    Z = Z * signal.get_window("hann", 1024, fftbins=True).sum()
    assert D.shape == Z.shape
    assert np.abs(np.abs(D) - np.abs(Z)).max() <= 1e-9 * np.abs(Z).max()


def test_oracle_features_shapes_ranges_and_trim():
    y = _utterance(2)
    mel, lin = FO.features(y, 22050, CFG)
    s, e = VO.trim_bounds(y.astype(np.float64), top_db=22.0)
    assert 3000 <= s <= 12000 and 30000 <= e <= 41000          # the near-silent lead / tail are cut
    frames = 1 + (e - s) // 256
    assert mel.shape == (80, frames // 4) and lin.shape == (513, 4 * (frames // 4))
    assert mel.dtype == np.float32 and lin.dtype == np.float32
    assert float(lin.max()) == 1.0 and 0.0 < float(mel.max()) <= 1.0 and float(mel.min()) >= 0.0
    log_cfg = dict(CFG, LOG_FEATURE=True)
    mel2, lin2 = FO.features(y, 22050, log_cfg)
    assert mel2.shape == mel.shape and 1e-8 <= float(mel2.min()) and float(lin2.max()) <= 1.0


def test_preemphasis_is_first_difference():
    y = np.arange(6, dtype=np.float32)
    assert np.allclose(FO.preemphasis(y, 0.97), [0, 1, 2 - 0.97, 3 - 1.94, 4 - 2.91, 5 - 3.88])


# ------------------------------------------------------------------------------------------- GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("log_feature", [False, True])
@pytest.mark.parametrize("seed,n", [(3, 40000), (4, 23456), (5, 9000)])
def test_wav_features_vs_oracle(seed, n, log_feature):
    import torch
    from spoofsv_b200.features import wav_features
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = dict(CFG, LOG_FEATURE=log_feature)
    y = _utterance(seed, n=n, lead=3000 + 100 * seed, tail=4000)
    omel, olin = FO.features(y, 22050, cfg)
    mel, lin = wav_features(y, 22050, cfg)
    mel, lin = mel.cpu().numpy(), lin.cpu().numpy()
    assert mel.shape == omel.shape and lin.shape == olin.shape          # same trim bounds, same frame count
    # fp32 FFT on the device vs float64 in the oracle; x ** 0.6 stretches errors of the quietest bins
    assert np.abs(lin - olin).max() <= 2e-3 and np.abs(mel - omel).max() <= 2e-3
    if log_feature:          # dB scale: a quiet bin's relative FFT error is an absolute error here
        assert np.abs(lin - olin).max() <= 5e-4 and np.abs(mel - omel).max() <= 5e-4
    else:
        strong = olin > 0.05
        assert np.abs(lin[strong] / olin[strong] - 1.0).max() <= 1e-4
        assert np.abs(mel / np.maximum(omel, 1e-3) - 1.0)[omel > 0.05].max() <= 1e-4


@pytest.mark.gpu
def test_feature_cache_layout_and_errors(tmp_path):
    import torch
    from scipy.io import wavfile
    from spoofsv_b200 import _lib
    from spoofsv_b200.features import cache_features, load_wav, wav_features
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    y = _utterance(6, n=20000)
    wav = tmp_path / "wav22" / "p225" / "p225_001.wav"
    wav.parent.mkdir(parents=True)
    wavfile.write(wav, 22050, (y * 32767).astype(np.int16))
    z, sr = load_wav(str(wav))
    assert sr == 22050 and z.dtype == np.float32 and abs(float(np.abs(z).max()) - float(np.abs(y).max())) < 1e-3
    mel, lin = cache_features(str(wav), str(tmp_path / "spec"), CFG)
    assert (tmp_path / "spec" / "p225" / "p225_001_mel.npy").exists()
    assert (tmp_path / "spec" / "p225" / "p225_001_lin.npy").exists()
    mel2, lin2 = cache_features(str(wav), str(tmp_path / "spec"), CFG)          # second call: from the cache
    assert np.array_equal(mel, mel2) and np.array_equal(lin, lin2)
    omel, olin = FO.features(z, sr, CFG)
    assert mel.shape == omel.shape and np.abs(mel - omel).max() <= 2e-3
    with pytest.raises(ValueError, match="too few"):
        wav_features(np.zeros(100, np.float32), 22050, CFG)
    lib = _lib.load()
    assert lib.ssv_spec_features(None, 513, 10, None, 80, 0, 0.6, 20.0, 100.0, 4, None, None, None, None) != 0
