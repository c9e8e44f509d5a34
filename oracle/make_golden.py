"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only)

What it does
  1. packs the reference's vendored inputs (spk_emb/*.npy, havard.txt) into tests/golden/;
  2. checks that the product's parameter containers reproduce the reference's state_dict
     (same keys, same order, bit-identical seeded weights);
  3. runs the reference modules (models/TTSModel.py: melSyn, SSRN, highwayConv) on seeded
     weights/inputs and stores their outputs as known-answer vectors;
  4. checks the oracle restatement (oracle/ttsmodel_oracle.py) against those outputs before
     writing them, so a stale oracle cannot produce its own goldens.
The reference drivers (generate_test_utterances.py, synthesize.py) need librosa / soundfile /
matplotlib, which are absent here; their loop is driven through the model API exactly as
generate_test_utterances.py:105-120 does.
"""
from __future__ import annotations

import shutil
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle import ttsmodel_oracle as O      # noqa: E402
from oracle import weights as W              # noqa: E402

warnings.filterwarnings("ignore")
GOLDEN = W.GOLDEN


def import_reference():
    sys.path.insert(0, str(REF))
    import importlib
    mod = importlib.import_module("models.TTSModel")
    assert str(REF) in mod.__file__, mod.__file__
    return mod


def ref_models(R, sd1, sd2):
    c = W.CFG
    r1 = R.melSyn(vocab_len=c["vocab_len"], condition=True, spkemb_dim=c["spkemb_dim"], textemb_dim=c["textemb_dim"],
                  freq_bins=c["freq_bins"], hidden_dim=c["hidden_dim"])
    r2 = R.SSRN(freq_bins=c["freq_bins"], output_bins=c["output_bins"], ssrn_dim=c["ssrn_dim"])
    r1.load_state_dict(sd1, strict=True)
    r2.load_state_dict(sd2, strict=True)
    return r1.eval(), r2.eval()


def ref_ar_loop(r1, textid, spk, n_frames):
    """generate_test_utterances.py:105-116 driven through the reference module."""
    B = textid.shape[0]
    init = torch.zeros((B, W.CFG["freq_bins"], 1))
    Y, A, pma, K, V = r1(melspec=init, textid=textid, spkemb=spk, pma=torch.zeros((B,)).long())
    traj = [pma.clone()]
    inputs = torch.cat((init, Y), dim=-1)
    for _ in range(n_frames - 1):
        Y, A, pma = r1(melspec=inputs, textid=None, spkemb=spk, K=K, V=V, A_last=A, pma=pma)
        traj.append(pma.clone())
        inputs = torch.cat((inputs, Y[:, :, -1:]), dim=-1)
    return Y, A, torch.stack(traj, 0), K, V


def close(a, b, tol, what):
    err = float((a - b).abs().max())
    assert err <= tol, f"{what}: oracle deviates from the reference by {err:.3e} (> {tol})"
    return err


def golden_train(R, ref_ids, emb):
    """Known answer for the train branch: the reference module in train() mode (TTSModel.py has no dropout, so
    train() only selects the branch), kaiming init + jittered LayerNorm affine, B = 3, T = 29."""
    k1, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
    r1, _ = ref_models(R, k1, k2)
    r1.train()
    ids = O.pad_text_ids([ref_ids[2], ref_ids[10], ref_ids[5]])
    spk = torch.from_numpy(emb[[0, 17, 55]])[:, :, None]
    g = torch.Generator().manual_seed(9)
    mel = torch.rand((3, W.CFG["freq_bins"], 29), generator=g)
    mel[:, :, 0] = 0                                           # teacher forcing: [0, mel_gt[:, :, :-1]]
    with torch.no_grad():
        Y, A = r1(mel, ids, spk)
        oY, oA = O.melsyn_train_forward(k1, mel, ids, spk)
    print("train oracle-vs-ref  Y %.2e  A %.2e" % (close(oY, Y, 2e-5, "train Y"), close(oA, A, 2e-5, "train A")))
    np.savez_compressed(GOLDEN / "train_seed7.npz", mel=mel.numpy(), textid=ids.numpy(), spk=spk.numpy(),
                        Y=Y.numpy(), A=A.numpy())


def golden_train_step(R, ref_ids, emb):
    """Known answers for one generator and one discriminator iteration of train/adversarial_wasserstein_gp.py:
    the unmodified reference melSyn (train mode) and melDisc (eval mode: its two dropouts off, everything else as
    in training), kaiming init + jittered LayerNorm affine, B = 3, T = 24.  The statements between the marks are
    the reference's own (:269-300 for G, :302-316 for D) with `device` / logging removed; the interpolation
    weights of the gradient penalty (reference: torch.rand(B)) are stored in the fixture."""
    import importlib
    D = importlib.import_module("models.discriminator")
    assert str(REF) in D.__file__
    k1, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
    model, _ = ref_models(R, k1, k2)
    model.train()
    torch.manual_seed(11)
    disc = D.melDisc(freq_bins=W.CFG["freq_bins"], disc_dim=128)
    for m in disc.modules():                                   # train/ordinary.py:16-19 init_weights
        if isinstance(m, torch.nn.Conv1d):
            torch.nn.init.kaiming_normal_(m.weight)
    disc.eval()
    ids = O.pad_text_ids([ref_ids[1], ref_ids[7], ref_ids[3]])
    spk_emb = torch.from_numpy(emb[[3, 40, 77]])[:, :, None]
    g = torch.Generator().manual_seed(21)
    mel_gt = torch.rand((3, W.CFG["freq_bins"], 24), generator=g) * 0.9 + 0.05
    coeff_b = torch.rand(3, generator=g)
    MAX_TEXT_LEN, MAX_FRAME_NUM, LAMBDA = 186, 325, 10
    import math
    gaw = torch.zeros((MAX_TEXT_LEN, MAX_FRAME_NUM))
    for a in range(MAX_TEXT_LEN):
        for b in range(MAX_FRAME_NUM):
            gaw[a, b] = 1 - math.exp(-(b / MAX_FRAME_NUM - a / MAX_TEXT_LEN) ** 2 / (2 * 0.2 * 0.2))
    F = torch.nn.functional
    text_id = ids
    B, C, T = mel_gt.shape
    # ---- G iteration (reference :269-300)
    spec_inputs = torch.cat((torch.zeros_like(mel_gt[:, :, :1]), mel_gt[:, :, :-1]), dim=-1)
    pred_mel_prob, att_mat = model(spec_inputs, text_id, spk_emb)
    disc_syn = disc(pred_mel_prob)
    loss_l1 = torch.mean(torch.abs(mel_gt - pred_mel_prob))
    loss_bin_div = torch.mean(-mel_gt * torch.log(pred_mel_prob + 1e-8) - (1 - mel_gt) * torch.log(1 - pred_mel_prob + 1e-8))
    att_aug = F.pad(att_mat, (0, MAX_FRAME_NUM - att_mat.size()[-1], 0, MAX_TEXT_LEN - att_mat.size()[-2]), value=-1)
    loss_att = torch.sum(torch.ne(att_aug, -1).float() * att_aug * gaw) / torch.sum(torch.ne(att_aug, -1).float())
    loss_disc = torch.mean(-disc_syn)
    loss = loss_l1 + loss_bin_div + loss_att + (loss_l1.item() + loss_bin_div.item() + loss_att.item()) / (abs(loss_disc.item())) * loss_disc
    loss.backward()
    g_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    g_terms = np.array([loss_l1.item(), loss_bin_div.item(), loss_att.item(), loss_disc.item(), loss.item()])
    disc.zero_grad()
    # ---- D iteration (reference :302-316)
    coeff = torch.stack(T * [torch.stack(C * [coeff_b], dim=1)], dim=2)
    input_mid = coeff * mel_gt.detach() + (1 - coeff) * pred_mel_prob.detach()
    input_mid.requires_grad = True
    output_mid = disc(input_mid)
    gradients = torch.autograd.grad(outputs=output_mid, inputs=input_mid, grad_outputs=torch.ones(output_mid.size()),
                                    retain_graph=True, create_graph=True)[0]
    loss_gp = torch.mean(LAMBDA * (torch.norm(gradients, p=2, dim=(1, 2)) - 1) ** 2)
    loss_gp.backward()
    disc_gt = disc(mel_gt.detach())
    disc_syn = disc(pred_mel_prob.detach())
    loss_D = torch.mean(disc_syn - disc_gt)
    loss_D.backward()
    d_terms = np.array([loss_gp.item(), -loss_D.item()])
    d_grads = {n: p.grad.detach().clone() for n, p in disc.named_parameters()}
    pick = ["text_encoder.hci1.hc2.conv.weight", "audio_encoder.hc2.ln1.weight", "audio_decoder.conv5.weight",
            "audio_encoder.fc1.weight", "text_encoder.textemb_layer.W.weight", "audio_decoder.hci.hc4.conv.bias"]
    out = {"mel_gt": mel_gt.numpy(), "textid": ids.numpy(), "spk": spk_emb.numpy(), "coeff": coeff_b.numpy(),
           "g_terms": g_terms, "d_terms": d_terms}
    for n in pick:                                             # big tensors: every 37th element keeps the fixture small
        gn = g_grads[n].numpy()
        out["ggrad/" + n] = gn.reshape(-1)[::37] if gn.size > 20000 else gn
    for n, v in disc.state_dict().items():
        out["disc/" + n] = v.numpy()
    for n in ("conv1.weight", "hc.conv.weight", "ln3.weight", "conv5.bias"):
        dn = d_grads[n].numpy()
        out["dgrad/" + n] = dn.reshape(-1)[::37] if dn.size > 20000 else dn
    np.savez_compressed(GOLDEN / "train_step_seed7.npz", **out)
    print("train step: G terms", g_terms, " D terms", d_terms)


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if "--train-step-only" in sys.argv:
        R = import_reference()
        torch.set_num_threads(8)
        lines = [ln.strip() for ln in (REF / "havard.txt").read_text().splitlines()]
        names = sorted(p.stem for p in (REF / "spk_emb").glob("*.npy"))
        emb = np.stack([np.load(REF / "spk_emb" / f"{n}.npy").astype(np.float32) for n in names])
        golden_train_step(R, [O.text2id(s) for s in lines], emb)
        return
    if "--train-only" in sys.argv:                             # add the train-branch vector without rewriting the rest
        R = import_reference()
        torch.set_num_threads(8)
        lines = [ln.strip() for ln in (REF / "havard.txt").read_text().splitlines()]
        names = sorted(p.stem for p in (REF / "spk_emb").glob("*.npy"))
        emb = np.stack([np.load(REF / "spk_emb" / f"{n}.npy").astype(np.float32) for n in names])
        golden_train(R, [O.text2id(s) for s in lines], emb)
        return
    R = import_reference()
    torch.set_num_threads(8)

    # ---- 1. vendored inputs
    names = sorted(p.stem for p in (REF / "spk_emb").glob("*.npy"))
    emb = np.stack([np.load(REF / "spk_emb" / f"{n}.npy").astype(np.float32) for n in names])
    np.savez_compressed(GOLDEN / "spk_emb.npz", names=np.array(names), emb=emb)
    shutil.copyfile(REF / "havard.txt", GOLDEN / "havard.txt")
    lines = [ln.strip() for ln in (REF / "havard.txt").read_text().splitlines()]

    # ---- 2. state_dict reproduction
    torch.manual_seed(0)
    c = W.CFG
    r1 = R.melSyn(vocab_len=c["vocab_len"], condition=True, spkemb_dim=c["spkemb_dim"], textemb_dim=c["textemb_dim"],
                  freq_bins=c["freq_bins"], hidden_dim=c["hidden_dim"])
    r2 = R.SSRN(freq_bins=c["freq_bins"], output_bins=c["output_bins"], ssrn_dim=c["ssrn_dim"])
    sd1, sd2 = W.state_dicts(0)
    for mine, ref in ((sd1, r1.state_dict()), (sd2, r2.state_dict())):
        assert list(mine.keys()) == list(ref.keys()), "state_dict keys/order differ from the reference"
        assert all(torch.equal(mine[k], ref[k]) for k in ref), "seeded weights differ from the reference"
    digests = {"t2m_seed0": W.sd_digest(sd1), "ssrn_seed0": W.sd_digest(sd2)}

    with torch.no_grad():
        # ---- 3a. text2id on the first 20 Harvard sentences (generate_test_utterances.py:60-73)
        char2idx = {ch: i for i, ch in enumerate(O.VOCABULARY)}
        char2idx['"'] = len(O.VOCABULARY) - 2
        import importlib.util
        # the reference text2id lives in files that import librosa; restate the call sites by exec'ing only the function
        src = (REF / "generate_test_utterances.py").read_text()
        start = src.index("def text2id"); end = src.index("# def plot_attention")
        ns = {"np": np}
        exec(compile(src[start:end], "ref_text2id", "exec"), ns)
        ref_ids = [ns["text2id"](lines[k], O.VOCABULARY, char2idx) for k in range(len(lines))]
        for k, rid in enumerate(ref_ids):
            assert np.array_equal(rid, O.text2id(lines[k])), f"text2id mismatch on line {k}"
        id_lens = np.array([r.shape[-1] for r in ref_ids], dtype=np.int64)

        # ---- 3b. cfg 1: seed 0, default init, Harvard #1, p225, 217 frames, SSRN
        r1, r2 = ref_models(R, sd1, sd2)
        ids1 = O.pad_text_ids([ref_ids[0]])
        spk = torch.from_numpy(emb[names.index("p225")])[None, :, None]
        T1 = 217
        Y, A, traj, K, V = ref_ar_loop(r1, ids1, spk, T1)
        lin = r2(Y)
        oY, oA, otraj, olin = O.synthesize(sd1, sd2, ids1, spk, T1)
        print("cfg1 oracle-vs-ref  Y %.2e  A %.2e  lin %.2e" % (
            close(oY, Y, 2e-5, "cfg1 Y"), close(oA, A, 2e-5, "cfg1 A"), close(olin, lin, 2e-5, "cfg1 lin")))
        assert torch.equal(otraj, traj), "cfg1 pma trajectory differs"
        iY, iA, itraj = O.ar_loop_incremental(sd1, ids1, spk, T1)
        print("cfg1 incremental-vs-ref  Y %.2e  A %.2e" % (close(iY, Y, 5e-5, "inc Y"), close(iA, A, 5e-5, "inc A")))
        assert torch.equal(itraj, traj)
        np.savez_compressed(
            GOLDEN / "cfg1_seed0.npz", textid=ids1.numpy(), spk=spk.numpy(), Y=Y.numpy(), A=A.numpy(),
            traj=traj.numpy(), K=K.numpy(), V=V.numpy(), lin_t4=lin[:, :, ::4].contiguous().numpy(),
            lin_first=lin[:, :, :16].contiguous().numpy())

        # ---- 3c. small batched case, kaiming init + jittered LayerNorm affine, ragged texts
        k1, k2 = W.state_dicts(7, init="kaiming", ln_jitter=True)
        digests["t2m_seed7_kaiming_jit"] = W.sd_digest(k1)
        digests["ssrn_seed7_kaiming_jit"] = W.sd_digest(k2)
        r1, r2 = ref_models(R, k1, k2)
        ids3 = O.pad_text_ids([ref_ids[2], ref_ids[10], ref_ids[5]])
        spk3 = torch.from_numpy(emb[[0, 17, 55]])[:, :, None]
        T3 = 40
        Y, A, traj, K, V = ref_ar_loop(r1, ids3, spk3, T3)
        lin = r2(Y)
        oY, oA, otraj, olin = O.synthesize(k1, k2, ids3, spk3, T3)
        print("small oracle-vs-ref  Y %.2e  A %.2e  lin %.2e" % (
            close(oY, Y, 2e-5, "small Y"), close(oA, A, 2e-5, "small A"), close(olin, lin, 2e-5, "small lin")))
        assert torch.equal(otraj, traj)
        iY, iA, itraj = O.ar_loop_incremental(k1, ids3, spk3, T3)
        close(iY, Y, 5e-5, "small inc Y"); close(iA, A, 5e-5, "small inc A")
        assert torch.equal(itraj, traj)
        np.savez_compressed(GOLDEN / "small_seed7.npz", textid=ids3.numpy(), spk=spk3.numpy(), Y=Y.numpy(),
                            A=A.numpy(), traj=traj.numpy(), K=K.numpy(), V=V.numpy(), lin=lin.numpy())

        # ---- 3d. SSRN alone (cfg 2 shape class, reduced): B=2, T=37 (odd, tile-straddling)
        g = torch.Generator().manual_seed(3)
        mel = torch.rand((2, 80, 37), generator=g)
        lin = r2(mel)
        close(O.ssrn(mel, k2), lin, 2e-5, "ssrn")
        np.savez_compressed(GOLDEN / "ssrn_seed7.npz", mel=mel.numpy(), lin=lin.numpy())

        # ---- 3e. single highwayConv, every (d, k, dilation, causal) class on the path
        cases = [(256, 3, 1, 0), (256, 3, 3, 0), (256, 3, 1, 1), (256, 3, 3, 1), (256, 3, 9, 1), (256, 3, 27, 1),
                 (512, 3, 1, 0), (512, 3, 3, 0), (512, 3, 9, 0), (512, 3, 27, 0), (512, 1, 1, 0)]
        out = {}
        for i, (d, k, dil, causal) in enumerate(cases):
            p = W.highway_params(d, k, 100 + i)
            hc = R.highwayConv(dimension=d, kernel_size=k, dilation=dil, causal=bool(causal)).eval()
            hc.load_state_dict(p, strict=True)
            g = torch.Generator().manual_seed(200 + i)
            x = torch.randn((2, d, 45), generator=g)
            y = hc(x)
            close(O.highway_conv(x, p, "", dil, bool(causal)) if False else
                  O.highway_conv(x, {("." + kk): v for kk, v in p.items()}, "", dil, bool(causal)), y, 1e-5, f"hc{i}")
            out[f"y{i}"] = y.numpy()
        out["cases"] = np.array(cases, dtype=np.int64)
        np.savez_compressed(GOLDEN / "highway_cases.npz", **out)

        # ---- 3f. train-mode (teacher-forced) forward of melSyn (models/TTSModel.py:263-273), ragged texts
        golden_train(R, ref_ids, emb)

    (GOLDEN / "digests.txt").write_text("".join(f"{k} {v}\n" for k, v in sorted(digests.items())))
    np.save(GOLDEN / "havard_id_lens.npy", id_lens)
    print("wrote", sorted(p.name for p in GOLDEN.iterdir()))


if __name__ == "__main__":
    main()
