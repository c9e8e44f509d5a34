"""CPU oracle for the SpoofSV Text2Mel + SSRN synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``spoofsv_b200/``)
imports this file; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

This is a functional restatement (state_dict in, tensors out) of the
reference's PyTorch modules, written against the reference semantics and
pinned to it by ``oracle/make_golden.py`` (which imports the unmodified
reference from /root/reference in the build container and writes
``tests/golden/*.npz``).  The arithmetic lives in PyTorch (reference pins
torch==1.2.0 in requirements.txt:5; this image has torch 2.11): conv1d,
layer_norm (eps 1e-5, biased variance), softmax, sigmoid, argmax (first max).

Reference anchors (``/root/reference/...``):
  models/TTSModel.py:25-35    textEmbedding       -> text_embedding
  models/TTSModel.py:63-84    highwayConv         -> highway_conv
  models/TTSModel.py:126-140  textEncoder         -> text_encoder
  models/TTSModel.py:166-184  audioEncoder        -> audio_encoder
  models/TTSModel.py:217-232  audioDecoder        -> audio_decoder
  models/TTSModel.py:275-300  melSyn.forward eval -> melsyn_eval_call
  models/TTSModel.py:342-362  SSRN.forward        -> ssrn
  generate_test_utterances.py:15-25,61-73   text2id + padding -> text2id / pad_text_ids
  generate_test_utterances.py:105-116       AR loop           -> ar_loop_reference
Tensors are channels-first (B, C, T) fp32 exactly like the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

VOCABULARY = "PE abcdefghijklmnopqrstuvwxyz-,.?'\""   # config.json:12
MASK_VALUE = float(-2 ** 32)                            # models/TTSModel.py:284


# --------------------------------------------------------------------------- text
def text2id(text: str, vocabulary: str = VOCABULARY) -> np.ndarray:
    """generate_test_utterances.py:15-25 / :60-61 -- lower-case, append 'E',
    drop characters outside the vocabulary, map '\"' onto the id of \"'\"."""
    lut = {ch: i for i, ch in enumerate(vocabulary)}
    lut['"'] = len(vocabulary) - 2
    ids = [lut[ch] for ch in (text.lower() + "E") if ch in vocabulary]
    return np.asarray(ids, dtype=np.int64)[None, :]


def pad_text_ids(id_rows: Sequence[np.ndarray]) -> torch.Tensor:
    """generate_test_utterances.py:68-73 -- right-pad with id 0 to the batch max,
    stack to (U, 1, N) int64."""
    n = max(r.shape[-1] for r in id_rows)
    out = torch.zeros((len(id_rows), 1, n), dtype=torch.int64)
    for u, r in enumerate(id_rows):
        out[u, 0, : r.shape[-1]] = torch.from_numpy(np.asarray(r).reshape(-1))
    return out


# --------------------------------------------------------------------------- layers
def _ln_ch(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    """LayerNorm over the channel axis of a (B, C, T) tensor."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    return F.layer_norm(x.transpose(1, 2), (w.numel(),), w, b, 1e-5).transpose(1, 2)


def _pw(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"])


def highway_conv(x: torch.Tensor, sd: SD, prefix: str, dilation: int, causal: bool) -> torch.Tensor:
    """models/TTSModel.py:63-84.  kernel size and width come from the weight."""
    w = sd[prefix + ".conv.weight"]
    d, k = w.shape[1], w.shape[2]
    pad = dilation * (k - 1) // 2
    if causal:
        xin = F.pad(x, (2 * pad, 0))
        h = F.conv1d(xin, w, sd[prefix + ".conv.bias"], dilation=dilation)
    else:
        h = F.conv1d(x, w, sd[prefix + ".conv.bias"], padding=pad, dilation=dilation)
    g = torch.sigmoid(_ln_ch(h[:, :d], sd, prefix + ".ln1"))
    c = _ln_ch(h[:, d:], sd, prefix + ".ln2")
    return g * c + (1 - g) * x


def _hci(x: torch.Tensor, sd: SD, prefix: str, causal: bool) -> torch.Tensor:
    """models/TTSModel.py:94-104: four highway convs, dilation 1,3,9,27."""
    for i, dil in enumerate((1, 3, 9, 27), start=1):
        x = highway_conv(x, sd, f"{prefix}.hc{i}", dil, causal)
    return x


def text_embedding(textid: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    """models/TTSModel.py:25-35: one-hot @ Linear == column gather + bias."""
    w, b = sd[prefix + ".W.weight"], sd[prefix + ".W.bias"]      # (E, vocab), (E,)
    ids = textid.long()[:, 0, :]                                  # (B, N)
    return (w.t()[ids] + b).transpose(1, 2)                       # (B, E, N)


def text_encoder(textid: torch.Tensor, sd: SD, p: str = "text_encoder") -> Tuple[torch.Tensor, torch.Tensor]:
    """models/TTSModel.py:126-140."""
    x = text_embedding(textid, sd, p + ".textemb_layer")
    x = _ln_ch(_pw(x, sd, p + ".conv1"), sd, p + ".ln1")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv2"), sd, p + ".ln2")
    x = _hci(x, sd, p + ".hci1", False)
    x = _hci(x, sd, p + ".hci2", False)
    x = highway_conv(x, sd, p + ".hc1", 1, False)
    x = highway_conv(x, sd, p + ".hc2", 1, False)
    x = highway_conv(x, sd, p + ".hc3", 1, False)
    x = highway_conv(x, sd, p + ".hc4", 1, False)
    h = x.shape[1] // 2
    return x[:, :h], x[:, h:]


def audio_encoder(mel: torch.Tensor, spk: torch.Tensor, sd: SD, p: str = "audio_encoder") -> torch.Tensor:
    """models/TTSModel.py:172-184 (condition branch). spk is (B, E, 1)."""
    s1 = F.linear(spk.transpose(1, 2), sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]).transpose(1, 2)
    s2 = F.linear(spk.transpose(1, 2), sd[p + ".fc2.weight"], sd[p + ".fc2.bias"]).transpose(1, 2)
    x = _ln_ch(_pw(mel, sd, p + ".conv1") + s1, sd, p + ".ln1")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv2"), sd, p + ".ln2")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv3") + s2, sd, p + ".ln3")
    x = _hci(x, sd, p + ".hci1", True)
    x = _hci(x, sd, p + ".hci2", True)
    x = highway_conv(x, sd, p + ".hc1", 3, True)
    x = highway_conv(x, sd, p + ".hc2", 3, True)
    return x


def audio_decoder(rq: torch.Tensor, sd: SD, p: str = "audio_decoder") -> torch.Tensor:
    """models/TTSModel.py:217-232."""
    x = _ln_ch(_pw(rq, sd, p + ".conv1"), sd, p + ".ln1")
    x = _hci(x, sd, p + ".hci", True)
    x = highway_conv(x, sd, p + ".hc1", 1, True)
    x = highway_conv(x, sd, p + ".hc2", 1, True)
    x = _ln_ch(_pw(x, sd, p + ".conv2"), sd, p + ".ln2")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv3"), sd, p + ".ln3")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv4"), sd, p + ".ln4")
    x = _ln_ch(_pw(F.relu(x), sd, p + ".conv5"), sd, p + ".ln5")
    return torch.sigmoid(x)


# --------------------------------------------------------------------------- Text2Mel train-mode (teacher-forced) forward
def melsyn_train_forward(sd: SD, melspec: torch.Tensor, textid: torch.Tensor, spkemb: torch.Tensor):
    """melSyn.forward in train() mode, models/TTSModel.py:263-273: full-sequence teacher-forced forward with
    unmasked attention over all characters.  Returns (Y_prob (B, F, T), A (B, N, T))."""
    K, V = text_encoder(textid, sd)
    Q = audio_encoder(melspec, spkemb, sd)
    A = F.softmax(torch.matmul(K.transpose(1, 2), Q) / math.sqrt(K.shape[1]), dim=1)
    R = torch.cat((torch.matmul(V, A), Q), dim=1)
    return audio_decoder(R, sd), A


# --------------------------------------------------------------------------- Text2Mel eval protocol
def melsyn_eval_call(sd: SD, melspec, textid, spkemb, K=None, V=None, A_last=None, pma=None):
    """One eval-mode call of melSyn.forward, models/TTSModel.py:275-300.

    Re-encodes the whole mel prefix, exactly like the reference."""
    T = melspec.shape[-1]
    B = melspec.shape[0]
    if T == 1:
        K, V = text_encoder(textid, sd)
    N = K.shape[-1]
    hidden = K.shape[1]
    Q = audio_encoder(melspec, spkemb, sd)
    A = torch.matmul(K.transpose(1, 2), Q) / math.sqrt(hidden)
    for b in range(B):
        p = int(pma[b])
        if p > 0:
            A[b, :p, -1] = MASK_VALUE
        if p + 2 < N - 1:
            A[b, p + 3:, -1] = MASK_VALUE
    A = F.softmax(A, dim=1)
    if T > 1:
        A = torch.cat((A_last, A[:, :, -1:]), dim=-1)
    max_att = torch.argmax(A, dim=1)[:, -1]
    R = torch.cat((torch.matmul(V, A), Q), dim=1)
    Y = audio_decoder(R, sd)
    if T == 1:
        return Y, A, max_att, K, V
    return Y, A, max_att


def ar_loop_reference(sd: SD, textid: torch.Tensor, spkemb: torch.Tensor, n_frames: int):
    """generate_test_utterances.py:105-116 with the frame count as a parameter.

    Returns Y (B, F, n_frames), A (B, N, n_frames), pma trajectory (n_frames, B)."""
    B = textid.shape[0]
    F_ = sd["audio_encoder.conv1.weight"].shape[1]
    init = torch.zeros((B, F_, 1))
    Y, A, pma, K, V = melsyn_eval_call(sd, init, textid, spkemb, pma=torch.zeros(B, dtype=torch.int64))
    traj = [pma.clone()]
    inputs = torch.cat((init, Y), dim=-1)
    for _ in range(n_frames - 1):
        Y, A, pma = melsyn_eval_call(sd, inputs, None, spkemb, K=K, V=V, A_last=A, pma=pma)
        traj.append(pma.clone())
        inputs = torch.cat((inputs, Y[:, :, -1:]), dim=-1)
    return Y, A, torch.stack(traj, 0)


# --------------------------------------------------------------------------- incremental model of the decode kernel
_ENC_HC = [("hci1.hc1", 1), ("hci1.hc2", 3), ("hci1.hc3", 9), ("hci1.hc4", 27),
           ("hci2.hc1", 1), ("hci2.hc2", 3), ("hci2.hc3", 9), ("hci2.hc4", 27),
           ("hc1", 3), ("hc2", 3)]
_DEC_HC = [("hci.hc1", 1), ("hci.hc2", 3), ("hci.hc3", 9), ("hci.hc4", 27), ("hc1", 1), ("hc2", 1)]


def _ln_vec(x, sd, name):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    return F.layer_norm(x, (w.numel(),), w, b, 1e-5)


def _hc_inc(hist: torch.Tensor, t: int, sd: SD, prefix: str, dil: int) -> torch.Tensor:
    """One causal highway conv evaluated at time t only. hist is the layer's
    input history (B, Tmax, d); rows < 0 are zero."""
    w, b = sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"]   # (2d, d, 3)
    d = w.shape[1]
    acc = b.unsqueeze(0).expand(hist.shape[0], -1).clone()
    for j in range(3):
        tt = t - (2 - j) * dil
        if tt >= 0:
            acc = acc + hist[:, tt, :] @ w[:, :, j].t()
    g = torch.sigmoid(_ln_vec(acc[:, :d], sd, prefix + ".ln1"))
    c = _ln_vec(acc[:, d:], sd, prefix + ".ln2")
    return g * c + (1 - g) * hist[:, t, :]


def ar_loop_incremental(sd: SD, textid: torch.Tensor, spkemb: torch.Tensor, n_frames: int):
    """O(T) restatement of the AR loop: per-layer causal-conv history instead of
    re-encoding the prefix.  Mathematically equal to
    ar_loop_reference; this is the CPU model of the CUDA decode kernel."""
    B = textid.shape[0]
    K, V = text_encoder(textid, sd)                    # (B, h, N)
    h, N = K.shape[1], K.shape[2]
    F_ = sd["audio_encoder.conv1.weight"].shape[1]
    e = spkemb[:, :, 0]
    pe, pd = "audio_encoder", "audio_decoder"
    s1 = F.linear(e, sd[pe + ".fc1.weight"], sd[pe + ".fc1.bias"])
    s2 = F.linear(e, sd[pe + ".fc2.weight"], sd[pe + ".fc2.bias"])
    enc_hist = [torch.zeros(B, n_frames, h) for _ in _ENC_HC]
    dec_hist = [torch.zeros(B, n_frames, h) for _ in _DEC_HC]
    Y = torch.zeros(B, F_, n_frames)
    A = torch.zeros(B, N, n_frames)
    pma = torch.zeros(B, dtype=torch.int64)
    traj = []
    x = torch.zeros(B, F_)
    pw = lambda v, name: F.linear(v, sd[name + ".weight"][:, :, 0], sd[name + ".bias"])
    for t in range(n_frames):
        u = _ln_vec(pw(x, pe + ".conv1") + s1, sd, pe + ".ln1")
        u = _ln_vec(pw(F.relu(u), pe + ".conv2"), sd, pe + ".ln2")
        u = _ln_vec(pw(F.relu(u), pe + ".conv3") + s2, sd, pe + ".ln3")
        for i, (name, dil) in enumerate(_ENC_HC):
            enc_hist[i][:, t] = u
            u = _hc_inc(enc_hist[i], t, sd, f"{pe}.{name}", dil)
        q = u
        logits = torch.einsum("bhn,bh->bn", K, q) / math.sqrt(h)
        idx = torch.arange(N)[None, :]
        lo = pma[:, None]
        masked = ((idx < lo) & (lo > 0)) | ((idx > lo + 2) & (lo + 2 < N - 1))
        logits = torch.where(masked, torch.full_like(logits, MASK_VALUE), logits)
        a = F.softmax(logits, dim=1)
        pma = torch.argmax(a, dim=1)
        traj.append(pma.clone())
        A[:, :, t] = a
        r = torch.einsum("bhn,bn->bh", V, a)
        u = _ln_vec(pw(torch.cat((r, q), 1), pd + ".conv1"), sd, pd + ".ln1")
        for i, (name, dil) in enumerate(_DEC_HC):
            dec_hist[i][:, t] = u
            u = _hc_inc(dec_hist[i], t, sd, f"{pd}.{name}", dil)
        u = _ln_vec(pw(u, pd + ".conv2"), sd, pd + ".ln2")
        u = _ln_vec(pw(F.relu(u), pd + ".conv3"), sd, pd + ".ln3")
        u = _ln_vec(pw(F.relu(u), pd + ".conv4"), sd, pd + ".ln4")
        u = _ln_vec(pw(F.relu(u), pd + ".conv5"), sd, pd + ".ln5")
        x = torch.sigmoid(u)
        Y[:, :, t] = x
    return Y, A, torch.stack(traj, 0)


# --------------------------------------------------------------------------- SSRN
def upsampling(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """models/TTSModel.py:313-317."""
    x = F.conv_transpose1d(x, sd[p + ".deconv.weight"], sd[p + ".deconv.bias"], stride=2)
    x = highway_conv(x, sd, p + ".hc1", 1, False)
    return highway_conv(x, sd, p + ".hc2", 3, False)


def ssrn(mel: torch.Tensor, sd: SD) -> torch.Tensor:
    """models/TTSModel.py:342-362. (B, F, T) -> (B, O, 4T)."""
    x = _ln_ch(_pw(mel, sd, "conv1"), sd, "ln1")
    x = highway_conv(x, sd, "hc1", 1, False)
    x = highway_conv(x, sd, "hc2", 3, False)
    x = upsampling(x, sd, "ups1")
    x = upsampling(x, sd, "ups2")
    x = _ln_ch(_pw(x, sd, "conv2"), sd, "ln2")
    x = highway_conv(x, sd, "hc3", 1, False)
    x = highway_conv(x, sd, "hc4", 1, False)
    x = _ln_ch(_pw(x, sd, "conv3"), sd, "ln3")
    x = _ln_ch(_pw(x, sd, "conv4"), sd, "ln4")
    x = _ln_ch(_pw(F.relu(x), sd, "conv5"), sd, "ln5")
    x = _ln_ch(_pw(F.relu(x), sd, "conv6"), sd, "ln6")
    return torch.sigmoid(x)


def synthesize(sd_t2m: SD, sd_ssrn: SD, textid, spkemb, n_frames: int, incremental: bool = False):
    """generate_test_utterances.py:105-120: AR loop then SSRN."""
    loop = ar_loop_incremental if incremental else ar_loop_reference
    Y, A, traj = loop(sd_t2m, textid, spkemb, n_frames)
    return Y, A, traj, ssrn(Y, sd_ssrn)
