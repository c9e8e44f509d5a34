"""main.py-compatible CLI (synthesize / adversarial training) on a miniature corpus laid out like the reference's."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _corpus(tmp_path, n_utts=6):
    from scipy.io import wavfile
    from oracle import weights as W
    from spoofsv_b200 import text as T
    from oracle.testsignals import utterance
    names, emb, lines = W.load_fixtures()
    root = tmp_path / "data"
    wavs, txts = [], []
    for s_i, spk in enumerate(("p225", "p226")):
        (root / "wav22" / spk).mkdir(parents=True)
        (root / "txt" / spk).mkdir(parents=True)
        (tmp_path / "spk_emb").mkdir(exist_ok=True)
        np.save(tmp_path / "spk_emb" / f"{spk}.npy", emb[names.index(spk)])
        for k in range(n_utts // 2):
            stem = f"{spk}_{k + 1:03d}"
            y = utterance(10 * s_i + k, n=9000 + 1500 * k, lead=1500, tail=1500)
            wavfile.write(root / "wav22" / spk / f"{stem}.wav", 22050, (y * 32767).astype(np.int16))
            (root / "txt" / spk / f"{stem}.txt").write_text(lines[3 * s_i + k][:24] + "\n")
            wavs.append(str(root / "wav22" / spk / f"{stem}.wav"))
            txts.append(str(root / "txt" / spk / f"{stem}.txt"))
    (root / "data_path" / "ordinary").mkdir(parents=True)
    for mode, sel in (("train", slice(0, n_utts)), ("validate", slice(0, 2)), ("synthesize", slice(1, 4))):
        (root / "data_path" / "ordinary" / f"wav.path.{mode}").write_text("\n".join(wavs[sel]) + "\n")
        (root / "data_path" / "ordinary" / f"txt.path.{mode}").write_text("\n".join(txts[sel]) + "\n")
    cfg = {"DATA_ROOT_DIR": str(root) + "/", "SPK_EMB_DIR": str(tmp_path / "spk_emb") + "/", "SRC_ROOT_DIR": str(tmp_path) + "/",
           "SPK_EMB_DIM": 200, "HIDDEN_DIM": 256, "TEXT_EMB_DIM": 128, "SSRN_DIM": 256, "DISC_DIM": 128,
           "VOCABULARY": T.DEFAULT_VOCABULARY, "MAX_TEXT_LEN": 186, "MAX_FRAME_NUM": 325, "SAMPLING_RATE": 22050,
           "PREEMPH": 0.97, "STFT": {"FFT_LENGTH": 1024, "HOP_LENGTH": 256}, "COARSE_MELSPEC": {"REDUCTION": 4, "FREQ_BINS": 80},
           "NORM_POWER": {"ANALYSIS": 0.6, "RECONSTRUCTION": 1.3}, "LOG_FEATURE": False, "MAX_DB": 100, "REF_DB": 20,
           "APPLY_DROPOUT": False, "MULTI_GPU": False, "BATCH_SIZE": 4, "MAX_EPOCHS": 50, "VAL_EVERY_ITER": 3,
           "ADAM": {"ALPHA": 2e-4, "BETA_1": 0.5, "BETA_2": 0.9, "EPSILON": 1e-6}, "RATIO": 2, "LAMBDA": 10,
           "INFERENCE_TEXT2MEL_MODEL": "absent", "INFERENCE_SSRN_MODEL": "absent"}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(cfg))
    return p, cfg


def test_main_synthesize_and_adversarial_training(tmp_path):
    import torch
    from scipy.io import wavfile
    from spoofsv_b200 import main as M
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg_path, cfg = _corpus(tmp_path)
    # synthesize: 3 items -> one batch, wavs S1..S3_B1, spectrogram cache filled
    assert M.main(["synthesize", "-C", str(cfg_path), "-T", "t0", "--random_init", "0", "--save_spectrogram", "--gl_iters", "4"]) == 0
    for k in (1, 2, 3):
        sr, w = wavfile.read(tmp_path / "samples" / "t0" / f"S{k}_B1.wav")
        assert sr == 22050 and w.dtype == np.float32 and abs(float(w.max()) - 0.75) < 1e-5
    assert (tmp_path / "spec" / "p225" / "p225_002_mel.npy").exists()
    # Text2Mel adversarial training: G at iterations 0, 3, 6 (RATIO 2), validation + checkpoint after iteration 3
    out = M.adversarial_train("train_text2mel", "conditional", cfg, str(tmp_path / "spec") + "/", None, "t1", max_iterations=7)
    assert out["iterations"] == 7 and np.isfinite(out["last_G"]) and np.isfinite(out["last_D"])
    ck_path = tmp_path / "checkpoints" / "conditional" / "adversarial" / "t1" / "text2mel_iteration_4.tar.pth"
    ck = torch.load(ck_path, map_location="cpu")
    assert {"epoch", "iteration", "model_state_dict", "disc_state_dict", "opt_state_dict_syn", "opt_state_dict_disc",
            "loss_val_log", "wd_log", "loss_train_log_syn", "loss_train_log_syn_onlyfromD", "loss_train_log_disc"} <= set(ck)
    assert len(ck["model_state_dict"]) == 214 and ck["iteration"] == 4
    assert (ck_path.parent / "text2mel_best_model.tar.pth").exists()
    # resume from it
    out2 = M.adversarial_train("train_text2mel", "conditional", cfg, str(tmp_path / "spec") + "/", str(ck_path), "t2", max_iterations=6)
    assert out2["iterations"] == 6
    # SSRN adversarial training
    out3 = M.adversarial_train("train_ssrn", "conditional", cfg, str(tmp_path / "spec") + "/", None, "t3", max_iterations=4)
    assert out3["iterations"] == 4 and np.isfinite(out3["last_G"]) and np.isfinite(out3["last_D"])
    assert (tmp_path / "checkpoints" / "conditional" / "adversarial" / "t3" / "ssrn_iteration_4.tar.pth").exists()
    # the trainer without a discriminator (train/ordinary.py), through the CLI
    assert M.main(["train_text2mel", "-C", str(cfg_path), "-T", "t4", "--save_spectrogram", "--max_iterations", "5"]) == 0
    ck = torch.load(tmp_path / "checkpoints" / "conditional" / "not_adversarial" / "t4" / "text2mel_iteration_4.tar.pth", map_location="cpu")
    assert set(ck) == {"epoch", "iteration", "model_state_dict", "optimizer_state_dict", "loss_val_log"}
    assert M.main(["train_ssrn", "-C", str(cfg_path), "-T", "t5", "--save_spectrogram", "--max_iterations", "2"]) == 0
