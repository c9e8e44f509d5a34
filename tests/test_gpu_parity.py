"""GPU parity: the CUDA path (through the drop-in modules -> C ABI) against the CPU oracle and the
committed known-answer vectors of the unmodified reference.

Tolerances (BASELINE.json north_star): FP32 path max-abs <= 1e-4 on mel / linear / attention with
identical alignment (pma) trajectories; BF16 path relative-L2 <= 2e-2.
"""
import numpy as np
import pytest
import torch

from oracle import ttsmodel_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _maxabs(a, b):
    return float((a.detach().cpu().float() - b.detach().cpu().float()).abs().max())


def _rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm())


# --------------------------------------------------------------------------- highwayConv
def test_highway_conv_all_classes_fp32(golden_dir):
    from spoofsv_b200.models import highwayConv
    z = np.load(golden_dir / "highway_cases.npz")
    for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
        hc = highwayConv(dimension=d, kernel_size=k, dilation=dil, causal=bool(causal))
        hc.load_state_dict(W.highway_params(d, k, 100 + i), strict=True)
        hc = hc.cuda().eval()
        x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
        y = hc(x.cuda())
        assert y.shape == x.shape
        assert _maxabs(y, _t(z[f"y{i}"])) <= FP32_TOL, (d, k, dil, causal)


@pytest.mark.parametrize("T", [1, 2, 31, 33, 130])
def test_highway_conv_ragged_lengths(T):
    """T smaller than the dilation reach, tile-straddling, and multi-tile."""
    from spoofsv_b200.models import highwayConv
    for (d, k, dil, causal) in [(256, 3, 27, 1), (256, 3, 3, 0), (512, 3, 9, 0)]:
        p = W.highway_params(d, k, 5)
        hc = highwayConv(d, k, dil, bool(causal))
        hc.load_state_dict(p, strict=True)
        x = torch.randn((3, d, T), generator=torch.Generator().manual_seed(T))
        want = O.highway_conv(x, {"." + kk: v for kk, v in p.items()}, "", dil, bool(causal))
        assert _maxabs(hc.cuda()(x.cuda()), want) <= FP32_TOL, (d, k, dil, causal, T)


def test_highway_conv_empty_and_bad_shape():
    from spoofsv_b200.models import highwayConv
    hc = highwayConv(256, 3, 1).cuda()
    assert hc(torch.zeros(0, 256, 5, device="cuda")).shape == (0, 256, 5)
    with pytest.raises(ValueError):
        hc(torch.zeros(1, 128, 5, device="cuda"))


# --------------------------------------------------------------------------- SSRN
def test_ssrn_golden_fp32(golden_dir, cuda_models_k):
    _, m2, _, _ = cuda_models_k
    z = np.load(golden_dir / "ssrn_seed7.npz")
    lin = m2(_t(z["mel"]).cuda())
    assert lin.shape == (2, 513, 148)
    assert _maxabs(lin, _t(z["lin"])) <= FP32_TOL


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma"])
def test_ssrn_both_fp32_arms(prec, golden_dir, cuda_models_k, cuda_models):
    """"fp32" = 12 of the 16 layers on the tensor cores with split operands (3xTF32), "fp32-ffma" = every layer on the
    CUDA cores; both hold the 1e-4 bar, on a short (two utterances per tile) and a ragged long input with odd batch."""
    _, mk, _, _ = cuda_models_k
    _, m2, _, sd2 = cuda_models
    old_k, old = mk.precision, m2.precision
    try:
        mk.precision = m2.precision = prec
        z = np.load(golden_dir / "ssrn_seed7.npz")
        assert _maxabs(mk(_t(z["mel"]).cuda()), _t(z["lin"])) <= FP32_TOL
        z = np.load(golden_dir / "cfg1_seed0.npz")
        lin = m2(_t(z["Y"]).cuda())
        assert _maxabs(lin[:, :, ::4], _t(z["lin_t4"])) <= FP32_TOL
        mel = torch.rand((3, 80, 131), generator=torch.Generator().manual_seed(3))
        got = m2(mel.cuda())
        assert _maxabs(got, O.ssrn(mel, sd2)) <= FP32_TOL
        one = m2(mel[2:3, :, :9].contiguous().cuda())
        assert _maxabs(one, O.ssrn(mel[2:3, :, :9], sd2)) <= FP32_TOL
    finally:
        mk.precision, m2.precision = old_k, old


def test_ssrn_noncontiguous_input_and_cfg1(golden_dir, cuda_models):
    _, m2, _, _ = cuda_models
    z = np.load(golden_dir / "cfg1_seed0.npz")
    Y = _t(z["Y"]).cuda()
    buf = torch.zeros((1, 80, 300), device="cuda")
    buf[:, :, :217] = Y
    lin = m2(buf[:, :, :217])                       # a view, like the decoder-owned Y buffer
    assert lin.shape == (1, 513, 868)
    assert _maxabs(lin[:, :, ::4], _t(z["lin_t4"])) <= FP32_TOL
    assert _maxabs(lin[:, :, :16], _t(z["lin_first"])) <= FP32_TOL


def test_ssrn_full_size_properties(cuda_models):
    """BASELINE config 2 size (32 x 217): range, determinism, batch independence."""
    _, m2, _, sd2 = cuda_models
    mel = torch.rand((32, 80, 217), generator=torch.Generator().manual_seed(0)).cuda()
    a = m2(mel)
    assert a.shape == (32, 513, 868)
    assert bool(torch.isfinite(a).all()) and float(a.min()) > 0 and float(a.max()) < 1
    assert torch.equal(a, m2(mel))
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(m2(mel[perm]), a[perm])
    want = O.ssrn(mel[5:6].cpu(), sd2)
    assert _maxabs(a[5:6], want) <= FP32_TOL


# --------------------------------------------------------------------------- text encoder
def test_text_encoder_golden(golden_dir, cuda_models_k):
    m1, _, _, _ = cuda_models_k
    z = np.load(golden_dir / "small_seed7.npz")
    K, V = m1.encode_text(_t(z["textid"]).cuda())
    assert _maxabs(K, _t(z["K"])) <= FP32_TOL and _maxabs(V, _t(z["V"])) <= FP32_TOL


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma"])
def test_text_encoder_both_fp32_arms(prec, golden_dir, cuda_models_k, cuda_models):
    """TextEnc on the tensor cores (tcgen05 kind::tf32, split operands: the default FP32 arm) and on the CUDA cores
    (FFMA, the cross-check) against the reference's K / V: the kaiming-init ragged batch, BASELINE config 1, and the
    bench batch (64 x 58, against the oracle)."""
    m1k, m1 = cuda_models_k[0], cuda_models[0]
    for m in (m1k, m1):
        m.precision = prec
    try:
        z = np.load(golden_dir / "small_seed7.npz")
        K, V = m1k.encode_text(_t(z["textid"]).cuda())
        assert _maxabs(K, _t(z["K"])) <= FP32_TOL and _maxabs(V, _t(z["V"])) <= FP32_TOL
        z = np.load(golden_dir / "cfg1_seed0.npz")
        K, V = m1.encode_text(_t(z["textid"]).cuda())
        assert _maxabs(K, _t(z["K"])) <= FP32_TOL
        ids = W.synthetic_text(64, 58, seed=11)
        with torch.no_grad():
            oK, oV = O.text_encoder(ids, cuda_models[2])
        K, V = m1.encode_text(ids.cuda())
        assert _maxabs(K, oK) <= FP32_TOL and _maxabs(V, oV) <= FP32_TOL
        # odd batch (the second utterance of the last 2 x 64-row tile is out of bounds), one character, 65 characters
        # (one utterance per 128-row tile), 130 (two tiles per utterance)
        for B, N in ((3, 58), (1, 1), (2, 65), (2, 130)):
            ids = W.synthetic_text(B, N, seed=B + N)
            with torch.no_grad():
                oK, oV = O.text_encoder(ids, cuda_models[2])
            K, V = m1.encode_text(ids.cuda())
            assert _maxabs(K, oK) <= FP32_TOL and _maxabs(V, oV) <= FP32_TOL, (B, N)
    finally:
        for m in (m1k, m1):
            m.precision = "fp32"


def test_highway_conv_ffma_arm_all_classes(golden_dir):
    """The CUDA-core FP32 arm (precision 'fp32-ffma'; the training backward's forward) stays pinned to the goldens."""
    from spoofsv_b200.models import highwayConv
    z = np.load(golden_dir / "highway_cases.npz")
    for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
        hc = highwayConv(dimension=d, kernel_size=k, dilation=dil, causal=bool(causal))
        hc.load_state_dict(W.highway_params(d, k, 100 + i), strict=True)
        hc = hc.cuda().eval()
        hc.precision = "fp32-ffma"
        x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
        assert _maxabs(hc(x.cuda()), _t(z[f"y{i}"])) <= FP32_TOL, (d, k, dil, causal)


def test_text_encoder_rejects_bad_ids(cuda_models):
    m1, _, _, _ = cuda_models
    with pytest.raises(ValueError):
        m1.encode_text(torch.full((1, 1, 4), 34, device="cuda"))
    with pytest.raises(ValueError):
        m1.encode_text(torch.zeros((1, 4), dtype=torch.int64, device="cuda"))
    # the decode path does not synchronise before launching: the device-side flag surfaces at the end of the batch
    bad = W.synthetic_text(2, 12, seed=2)
    bad[1, 0, 3] = 34
    with pytest.raises(ValueError, match="text id"):
        m1.synthesize(bad.cuda(), torch.full((2, 200, 1), 0.05, device="cuda"), 4)
    ok = W.synthetic_text(2, 12, seed=2)
    Y, _, _, _, _ = m1.synthesize(ok.cuda(), torch.full((2, 200, 1), 0.05, device="cuda"), 4)     # flag was cleared
    assert bool(torch.isfinite(Y).all())


# --------------------------------------------------------------------------- decode
def test_decode_cfg1_217_frames(golden_dir, cuda_models):
    """BASELINE config 1: identical alignment trajectory over 217 frames, mel within 1e-4."""
    m1, m2, _, _ = cuda_models
    z = np.load(golden_dir / "cfg1_seed0.npz")
    Y, A, traj, K, V = m1.synthesize(_t(z["textid"]).cuda(), _t(z["spk"]).cuda(), 217)
    assert np.array_equal(traj.cpu().numpy(), z["traj"])
    assert _maxabs(K, _t(z["K"])) <= FP32_TOL
    assert _maxabs(Y, _t(z["Y"])) <= FP32_TOL
    assert _maxabs(A, _t(z["A"])) <= FP32_TOL
    lin = m2(Y)
    assert _maxabs(lin[:, :, ::4], _t(z["lin_t4"])) <= FP32_TOL


def test_decode_small_batch_run_and_protocol(golden_dir, cuda_models_k):
    """Ragged batch of 3; the one-launch run and the reference's call-per-frame protocol agree."""
    m1, m2, _, _ = cuda_models_k
    z = np.load(golden_dir / "small_seed7.npz")
    ids, spk = _t(z["textid"]).cuda(), _t(z["spk"]).cuda()
    T = z["Y"].shape[-1]
    Y, A, traj, _, _ = m1.synthesize(ids, spk, T)
    assert np.array_equal(traj.cpu().numpy(), z["traj"])
    assert _maxabs(Y, _t(z["Y"])) <= FP32_TOL and _maxabs(A, _t(z["A"])) <= FP32_TOL
    assert _maxabs(m2(Y), _t(z["lin"])) <= FP32_TOL
    Y_run, A_run = Y.clone(), A.clone()

    # generate_test_utterances.py:105-116, verbatim call pattern
    init = torch.zeros((3, 80, 1), device="cuda")
    Yp, Ap, pma, K, V = m1(melspec=init, textid=ids, spkemb=spk, pma=torch.zeros((3,), device="cuda").long())
    assert Yp.shape == (3, 80, 1) and Ap.shape == (3, ids.shape[-1], 1) and pma.dtype == torch.int64
    inputs = torch.cat((init, Yp), dim=-1)
    for _ in range(T - 1):
        Yp, Ap, pma = m1(melspec=inputs, textid=None, spkemb=spk, K=K, V=V, A_last=Ap, pma=pma)
        inputs = torch.cat((inputs, Yp[:, :, -1:]), dim=-1)
    m1.check()
    assert Yp.shape == (3, 80, T) and Ap.shape == (3, ids.shape[-1], T)
    assert torch.equal(Yp, Y_run) and torch.equal(Ap, A_run)
    assert np.array_equal(pma.cpu().numpy(), z["traj"][-1])


def test_decode_protocol_violation_raises(cuda_models):
    m1, _, _, _ = cuda_models
    ids = W.synthetic_text(2, 12, seed=2).cuda()
    spk = torch.full((2, 200, 1), 0.05, device="cuda")
    m1(melspec=torch.zeros((2, 80, 1), device="cuda"), textid=ids, spkemb=spk, pma=torch.zeros(2, device="cuda").long())
    with pytest.raises(RuntimeError, match="one frame per call"):
        m1(melspec=torch.zeros((2, 80, 5), device="cuda"), textid=None, spkemb=spk, pma=torch.zeros(2, device="cuda").long())


@pytest.mark.parametrize("B", [1, 2, 4, 5, 16, 17, 64])
def test_decode_batch_sizes_vs_oracle(B, cuda_models):
    """Every row-group configuration of the decode kernel (1 / 4 / 16 rows, 1..4 groups), N = 58."""
    m1, _, sd1, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(B, 58, seed=B)
    spk = torch.from_numpy(emb[:B].copy())[:, :, None]
    T = 30
    Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
    with torch.no_grad():
        oY, oA, otraj = O.ar_loop_incremental(sd1, ids, spk, T)
    assert np.array_equal(traj.cpu().numpy(), otraj.numpy())
    assert _maxabs(Y, oY) <= FP32_TOL and _maxabs(A, oA) <= FP32_TOL


@pytest.mark.parametrize("B,R,Wp,F", [(3, 2, None, None), (5, 4, None, None), (7, 1, None, None), (41, None, None, None),
                                      (67, 4, None, None), (130, None, None, None), (9, 1, 1, None), (6, 1, 2, None),
                                      (26, None, None, None), (13, 2, 1, None), (21, 1, 4, None),
                                      (1, 1, 4, 8), (7, 1, 2, 8), (9, 2, 1, 8), (5, 2, 2, 8), (3, 2, 4, 8), (6, 4, 1, 8),
                                      (9, 4, 2, 8), (67, 2, 2, 8), (131, 4, 2, 8), (70, 2, 1, 8), (200, None, None, None)])
def test_decode_micro_batch_shapes_vs_oracle(B, R, Wp, F, cuda_models):
    """Weight-stationary pipeline, every front-end shape: rows per micro-batch R = 1 / 2 / 4, warps per row
    W = 1 / 2 / 4 and F = 4 / 8 front-end warps per CTA (forced through ssv_decoder_set_plan and chosen by ws_plan),
    F / (R W) micro-batches in flight, ragged last micro-batch, micro-batch count padded to the in-flight count
    (dead micro-batches), more micro-batches than pipeline stages."""
    m1, _, sd1, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(B, 40, seed=100 + B)
    spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None]
    T = 64 if B <= 9 else 12           # small batches: long enough for the dilation-27 taps to leave the zero region
    m1.decode_plan = (R or 0, Wp or 0, F or 0)
    try:
        Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
    finally:
        m1.decode_plan = None
    with torch.no_grad():
        oY, oA, otraj = O.ar_loop_incremental(sd1, ids, spk, T)
    assert np.array_equal(traj.cpu().numpy(), otraj.numpy())
    assert _maxabs(Y, oY) <= FP32_TOL and _maxabs(A, oA) <= FP32_TOL


def test_decode_teacher_forced_frames(cuda_models):
    """The step API takes the newest column of the caller's melspec (not its own y) and the caller's pma."""
    m1, _, sd1, _ = cuda_models
    ids = W.synthetic_text(2, 20, seed=3)
    spk = torch.full((2, 200, 1), 0.06)
    mel = torch.rand((2, 80, 6), generator=torch.Generator().manual_seed(4))
    mel[:, :, 0] = 0
    with torch.no_grad():
        K, V = O.text_encoder(ids, sd1)
        pma = torch.zeros(2, dtype=torch.int64)
        A_last = None
        want = []
        for t in range(1, 7):
            if t == 1:
                Y, A_last, pma, _, _ = O.melsyn_eval_call(sd1, mel[:, :, :1], ids, spk, pma=pma)
            else:
                Y, A_last, pma = O.melsyn_eval_call(sd1, mel[:, :, :t], None, spk, K=K, V=V, A_last=A_last, pma=pma)
            want.append((Y.clone(), pma.clone()))
    melc, spkc = mel.cuda(), spk.cuda()
    pm = torch.zeros(2, dtype=torch.int64, device="cuda")
    for t in range(1, 7):
        if t == 1:
            Y, A, pm, Kc, Vc = m1(melspec=melc[:, :, :1], textid=ids.cuda(), spkemb=spkc, pma=pm)
        else:
            Y, A, pm = m1(melspec=melc[:, :, :t], textid=None, spkemb=spkc, K=Kc, V=Vc, A_last=A, pma=pm)
        assert _maxabs(Y, want[t - 1][0]) <= FP32_TOL
        assert torch.equal(pm.cpu(), want[t - 1][1])
    m1.check()


def test_decode_continuation_equals_single_run(cuda_models):
    """run(T) == step-wise continuation; chunked launches leave the same state."""
    m1, _, _, _ = cuda_models
    from spoofsv_b200 import _lib
    import ctypes as C
    ids = W.synthetic_text(4, 33, seed=5).cuda()
    spk = torch.full((4, 200, 1), 0.05, device="cuda")
    Y, A, traj, K, V = m1.synthesize(ids, spk, 60)
    Y, A, traj = Y.clone(), A.clone(), traj.clone()
    dec = m1._begin(K, V, spk, 60)
    lib = _lib.load()
    for n in (1, 7, 30, 22):
        _lib.check(lib.ssv_decoder_run(dec, n, _lib.current_stream_ptr()))
    _lib.check(lib.ssv_decoder_check(dec, _lib.current_stream_ptr()))
    st = m1._state
    assert torch.equal(st["Y"], Y) and torch.equal(st["A"], A) and torch.equal(st["traj"], traj)
    with pytest.raises(_lib.SsvError, match="capacity"):
        _lib.check(lib.ssv_decoder_run(dec, 1, _lib.current_stream_ptr()))


def test_full_size_decode_properties(cuda_models):
    """BASELINE config 3 size (B=64, N=58, 217 frames): determinism, batch independence, ranges,
    monotone window constraint of the alignment."""
    m1, _, _, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(64, 58, seed=11).cuda()
    spk = torch.from_numpy(emb[:64].copy())[:, :, None].cuda()
    Y, A, traj, _, _ = m1.synthesize(ids, spk, 217)
    Y, A, traj = Y.clone(), A.clone(), traj.clone()
    assert bool(torch.isfinite(Y).all()) and float(Y.min()) > 0 and float(Y.max()) < 1
    assert torch.allclose(A.sum(1), torch.ones_like(A.sum(1)), atol=1e-5)
    assert int((A > 0).sum(1).max()) <= 3
    step = traj[1:] - traj[:-1]
    assert int(step.min()) >= 0 and int(step.max()) <= 2          # argmax stays inside [pma, pma+2]
    Y2, A2, traj2, _, _ = m1.synthesize(ids, spk, 217)
    assert torch.equal(Y2, Y) and torch.equal(traj2, traj)
    sub = [3, 40, 63]
    Ys, As, trajs, _, _ = m1.synthesize(ids[sub], spk[sub], 217)
    assert torch.equal(trajs, traj[:, sub])
    assert _maxabs(Ys, Y[sub]) <= 1e-5


@pytest.mark.parametrize("B,T", [(64, 217), (41, 120), (130, 120)])
def test_decode_bench_plans_vs_oracle_all_rows(B, T, cuda_models):
    """The plans ws_plan picks for 33 <= B <= 128 (R = 2, W = 1: the kernel bench.py times) and beyond, at N = 58
    and frame counts well past the dilation-27 reach (t >= 54), so the shared-memory tap image and the global history
    rings of the dilation-27 layers carry non-zero history: every row of Y, A and the alignment trajectory against
    the oracle's incremental loop (BASELINE config 3 at B = 64: the bench workload itself)."""
    m1, _, sd1, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(B, 58, seed=11)
    spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None]
    Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
    with torch.no_grad():
        oY, oA, otraj = O.ar_loop_incremental(sd1, ids, spk, T)
    assert np.array_equal(traj.cpu().numpy(), otraj.numpy())
    assert _maxabs(Y, oY) <= FP32_TOL and _maxabs(A, oA) <= FP32_TOL


def test_reference_driver_maximum_sizes_vs_oracle(cuda_models):
    """The stock generate_test_utterances.py loop at its configured maxima: 20 sentences per speaker padded to
    MAX_TEXT_LEN = 186, MAX_FRAME_NUM + 1 = 326 frames (config.json:13-14, generate_test_utterances.py:105-116);
    two of the rows are checked against the oracle's incremental loop, all of them for the window constraint."""
    m1, _, sd1, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    B, N, T = 20, 186, 326
    ids = W.synthetic_text(B, N, seed=77)
    spk = torch.from_numpy(emb[40:40 + B].copy())[:, :, None]
    Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
    assert tuple(Y.shape) == (B, 80, T) and tuple(A.shape) == (B, N, T)
    step = traj[1:] - traj[:-1]
    assert int(step.min()) >= 0 and int(step.max()) <= 2 and int(traj.max()) <= N - 1
    sub = [0, B - 1]
    with torch.no_grad():
        oY, oA, otraj = O.ar_loop_incremental(sd1, ids[sub], spk[sub], T)
    assert np.array_equal(traj[:, sub].cpu().numpy(), otraj.numpy())
    assert _maxabs(Y[sub], oY) <= FP32_TOL and _maxabs(A[sub], oA) <= FP32_TOL


# --------------------------------------------------------------------------- host-buffer entry point
def test_synthesize_host_matches_module_path(cuda_models):
    from spoofsv_b200.synth import Synthesizer
    m1, m2, _, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(3, 25, seed=8)
    spk = emb[:3].copy()
    syn = Synthesizer(m1, m2)
    out = syn.synthesize_host(ids[:, 0, :].numpy(), spk, 20, want_mel=True, want_att=True)
    Y, A, traj, _, _ = m1.synthesize(ids.cuda(), torch.from_numpy(spk)[:, :, None].cuda(), 20)
    lin = m2(Y)
    assert np.array_equal(out["traj"], traj.cpu().numpy())
    assert np.array_equal(out["mel"], Y.cpu().numpy()) and np.array_equal(out["lin"], lin.cpu().numpy())
    assert np.array_equal(out["att"], A.cpu().numpy())


# --------------------------------------------------------------------------- BF16 tensor-core (tcgen05) path
BF16_TOL = 2e-2   # relative L2 (BASELINE.json north_star)


def test_highway_conv_all_classes_bf16(golden_dir):
    from spoofsv_b200.models import highwayConv
    z = np.load(golden_dir / "highway_cases.npz")
    for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
        hc = highwayConv(dimension=d, kernel_size=k, dilation=dil, causal=bool(causal))
        hc.load_state_dict(W.highway_params(d, k, 100 + i), strict=True)
        hc = hc.cuda().eval()
        hc.precision = "bf16"
        x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
        y = hc(x.cuda())
        assert _rel_l2(y, _t(z[f"y{i}"])) <= BF16_TOL, (d, k, dil, causal)


@pytest.mark.parametrize("T", [1, 127, 128, 129, 300])
def test_highway_conv_bf16_tile_edges(T):
    """Rows beyond T inside the last 128-row tile, taps reaching outside [0, T) (TMA zero fill)."""
    from spoofsv_b200.models import highwayConv
    for (d, k, dil, causal) in [(256, 3, 27, 1), (256, 3, 3, 0), (512, 3, 1, 0)]:
        p = W.highway_params(d, k, 6)
        hc = highwayConv(d, k, dil, bool(causal))
        hc.load_state_dict(p, strict=True)
        hc.precision = "bf16"
        x = torch.randn((3, d, T), generator=torch.Generator().manual_seed(T))
        want = O.highway_conv(x, {"." + kk: v for kk, v in p.items()}, "", dil, bool(causal))
        assert _rel_l2(hc.cuda()(x.cuda()), want) <= BF16_TOL, (d, k, dil, causal, T)


def test_ssrn_bf16(golden_dir, cuda_models_k, cuda_models):
    _, m2, _, _ = cuda_models_k
    z = np.load(golden_dir / "ssrn_seed7.npz")
    m2.precision = "bf16"
    try:
        lin = m2(_t(z["mel"]).cuda())
    finally:
        m2.precision = "fp32"
    assert lin.shape == (2, 513, 148)
    assert _rel_l2(lin, _t(z["lin"])) <= BF16_TOL
    _, m2b, _, sd2 = cuda_models
    mel = torch.rand((32, 80, 217), generator=torch.Generator().manual_seed(0)).cuda()
    m2b.precision = "bf16"
    try:
        a = m2b(mel)
        assert torch.equal(a, m2b(mel))
    finally:
        m2b.precision = "fp32"
    # BASELINE config 2 (32 x 217 -> 32 x 513 x 868): the tcgen05 arm against the ORACLE on every utterance
    with torch.no_grad():
        ref = O.ssrn(mel.cpu(), sd2)
    assert _rel_l2(a, ref) <= BF16_TOL
    for b in (0, 13, 31):
        assert _rel_l2(a[b], ref[b]) <= BF16_TOL
    assert float(a.min()) >= 0 and float(a.max()) <= 1


# --------------------------------------------------------------------------- corpus driver
def test_corpus_driver_matches_per_speaker_reference_loop(tmp_path, write_driver_cfg):
    """generate_test_utterances driver (K/V hoisted out of the speaker loop, units batched across speakers)
    == the reference's per-speaker AR loop + SSRN, as restated by the oracle."""
    import json
    from spoofsv_b200 import generate_test_utterances as G
    path, cfg = write_driver_cfg(tmp_path, n_lines=3, speakers=("p225", "p226"))
    args = G.build_parser().parse_args(["-C", str(path), "-T", "t", "--eval_utt_num", "3", "--batch", "4",
                                        "--ssrn_precision", "fp32", "--random_init", "0",
                                        "--save_spectrogram", str(tmp_path / "out")])
    got = {}

    def on_batch(group, lin, state):
        for u, spec, y in zip(group, lin, state["Y"].cpu()):
            got[(u.speaker, u.sentence)] = (spec.copy(), y.clone())

    stats = G.run(args, on_batch)
    assert stats["utterances"] == 6 and stats["frames"] == 10
    sd1, sd2 = W.state_dicts(0)
    names, emb, lines = W.load_fixtures()
    ids = O.pad_text_ids([O.text2id(s) for s in lines[:3]])
    for si, spk in enumerate(("p225", "p226")):
        e = torch.from_numpy(emb[names.index(spk)])[None, :, None].repeat(3, 1, 1)
        with torch.no_grad():
            oY, oA, otraj, olin = O.synthesize(sd1, sd2, ids, e, 10)
        for k in range(3):
            spec, y = got[(si, k)]
            assert _maxabs(y, oY[k]) <= FP32_TOL and _maxabs(_t(spec), olin[k]) <= FP32_TOL
            saved = np.load(tmp_path / "out" / f"s{spk[1:]}" / f"s{spk[1:]}_{k + 1:03d}.npy")
            assert np.array_equal(saved, spec)


def test_pipelined_submit_wait_matches_sync(cuda_models):
    """ssv_synthesize_host_submit / _wait with two batches in flight == the synchronous call, batch by batch."""
    from spoofsv_b200 import _lib
    from spoofsv_b200.synth import Synthesizer
    m1, m2, _, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    syn = Synthesizer(m1, m2)
    batches = [(W.synthetic_text(4, 25, seed=30 + i).numpy()[:, 0, :], emb[4 * i:4 * i + 4].copy()) for i in range(3)]
    want = [{k: v.copy() for k, v in syn.synthesize_host(ids, spk, 12, want_mel=True).items()} for ids, spk in batches]
    pend, got = None, []
    for ids, spk in batches:
        cur = syn.submit(ids, spk, 12, want_mel=True)
        if pend is not None:
            got.append({k: v.copy() for k, v in syn.collect(pend).items()})
        pend = cur
    got.append({k: v.copy() for k, v in syn.collect(pend).items()})
    for w, g in zip(want, got):
        assert np.array_equal(w["lin"], g["lin"]) and np.array_equal(w["mel"], g["mel"]) and np.array_equal(w["traj"], g["traj"])
    # protocol errors: a third submit while two are in flight, waiting twice for the same ticket
    a = syn.submit(*batches[0], 12)
    b = syn.submit(*batches[1], 12)
    with pytest.raises(ValueError, match="in flight"):
        syn.submit(*batches[2], 12)
    syn.collect(a); syn.collect(b)
    with pytest.raises(ValueError, match="no batch in flight"):
        syn.collect(a)


def test_host_result_as_bf16(cuda_models):
    """ssv_decoder_set_lin_output: the half-size device->host result equals the fp32 one rounded to bfloat16."""
    from spoofsv_b200.synth import Synthesizer
    m1, m2, _, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(3, 25, seed=8).numpy()[:, 0, :]
    spk = emb[:3].copy()
    want = Synthesizer(m1, m2).synthesize_host(ids, spk, 12)["lin"].copy()
    got = Synthesizer(m1, m2, lin_dtype="bf16").synthesize_host(ids, spk, 12)["lin"]
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == want.shape
    assert torch.equal(got, torch.from_numpy(want).to(torch.bfloat16))
    back = Synthesizer(m1, m2).synthesize_host(ids, spk, 12)["lin"]          # and back to fp32 on the same decoder
    assert np.array_equal(back, want)


def test_corpus_driver_writes_wavs(tmp_path, write_driver_cfg):
    """--save_wav: the reference's file naming and sample format (float32 wav at SAMPLING_RATE, peak 0.75, <= 9 s)."""
    import json
    from scipy.io import wavfile
    from spoofsv_b200 import generate_test_utterances as G
    path, cfg = write_driver_cfg(tmp_path, n_lines=2, speakers=("p225",))
    cfg.update({"NORM_POWER": {"ANALYSIS": 0.6, "RECONSTRUCTION": 1.3}, "PREEMPH": 0.97, "SAMPLING_RATE": 22050,
                "LOG_FEATURE": False, "MAX_FRAME_NUM": 19})
    path.write_text(json.dumps(cfg))
    args = G.build_parser().parse_args(["-C", str(path), "-T", "t", "--eval_utt_num", "2", "--random_init", "0",
                                        "--save_wav", str(tmp_path / "wav"), "--gl_iters", "8"])
    assert G.run(args)["utterances"] == 2
    for k in (1, 2):
        sr, w = wavfile.read(tmp_path / "wav" / "s225" / f"s225_{k:03d}.wav")
        assert sr == 22050 and w.dtype == np.float32 and 0 < len(w) <= 9 * 22050
        assert abs(float(w.max()) - 0.75) < 1e-6 and np.isfinite(w).all()


# --------------------------------------------------------------------------- train branch (teacher-forced forward)
def test_train_mode_forward_golden(golden_dir, cuda_models_k):
    """melSyn.forward in train() mode (models/TTSModel.py:263-273) against the reference's own output."""
    m1, _, sd1, _ = cuda_models_k
    z = np.load(golden_dir / "train_seed7.npz")
    m1.train()
    try:
        Y, A = m1(_t(z["mel"]).cuda(), _t(z["textid"]).cuda(), _t(z["spk"]).cuda())      # positional, as the training loops call it
    finally:
        m1.eval()
    assert Y.shape == z["Y"].shape and A.shape == z["A"].shape
    assert _maxabs(Y, _t(z["Y"])) <= FP32_TOL and _maxabs(A, _t(z["A"])) <= FP32_TOL
    assert torch.allclose(A.sum(1), torch.ones_like(A.sum(1)), atol=1e-5)


def test_train_mode_forward_config5_shape_vs_oracle(cuda_models):
    """BASELINE config 5 shape (B = 32, N = 64, T = 217) against the oracle; teacher forcing makes the train forward
    and the incremental decode agree on the first frame."""
    m1, _, sd1, _ = cuda_models
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(32, 64, seed=41)
    spk = torch.from_numpy(emb[:32].copy())[:, :, None]
    mel = torch.rand((32, 80, 217), generator=torch.Generator().manual_seed(42))
    mel[:, :, 0] = 0
    m1.train()
    try:
        Y, A = m1(mel.cuda(), ids.cuda(), spk.cuda())
    finally:
        m1.eval()
    with torch.no_grad():
        oY, oA = O.melsyn_train_forward(sd1, mel, ids, spk)
    assert _maxabs(Y, oY) <= FP32_TOL and _maxabs(A, oA) <= FP32_TOL


# --------------------------------------------------------------------------- training step, first building block
@pytest.mark.parametrize("d,k,dil,causal,B,T", [(256, 3, 1, True, 2, 37), (256, 3, 27, True, 3, 70), (256, 3, 3, False, 2, 45),
                                                  (512, 3, 9, False, 2, 58), (512, 1, 1, False, 3, 33), (256, 3, 1, False, 5, 130)])
@pytest.mark.parametrize("save_h", [True, False])
def test_highway_conv_backward_vs_autograd(d, k, dil, causal, B, T, save_h, monkeypatch):
    """ssv_highway_conv_bwd against torch.autograd through the oracle's restatement of highwayConv.forward
    (models/TTSModel.py:63-84): input, conv, bias and LayerNorm gradients, few rows and many, every tap layout;
    with H saved by the training-time forward and with H recomputed in the backward pass."""
    from spoofsv_b200.models import TTSModel as TM
    from spoofsv_b200.models.TTSModel import highwayConv
    monkeypatch.setattr(TM._HighwayConvFn, "save_h", save_h)
    torch.manual_seed(d + k + dil)
    hc = highwayConv(d, k, dil, causal=causal).cuda()
    with torch.no_grad():
        for p in (hc.ln1.weight, hc.ln2.weight):
            p.add_(0.2 * torch.randn_like(p))
        for p in (hc.ln1.bias, hc.ln2.bias):
            p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, d, T, device="cuda", requires_grad=True)
    gy = torch.randn(B, d, T, device="cuda")
    y = hc(x)
    assert y.requires_grad
    y.backward(gy)
    got = [x.grad] + [p.grad for p in (hc.conv.weight, hc.conv.bias, hc.ln1.weight, hc.ln1.bias, hc.ln2.weight, hc.ln2.bias)]
    # oracle: the same computation in torch ops on the CPU, float64, differentiated by autograd
    sd = {"h." + n: p.detach().cpu().double().requires_grad_(True) for n, p in hc.state_dict().items()}
    xo = x.detach().cpu().double().requires_grad_(True)
    yo = O.highway_conv(xo, sd, "h", dil, causal)
    assert _maxabs(y, yo.float()) <= FP32_TOL
    yo.backward(gy.cpu().double())
    want = [xo.grad] + [sd["h." + n].grad for n in ("conv.weight", "conv.bias", "ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias")]
    for g, w in zip(got, want):
        assert g.shape == w.shape
        scale = float(w.abs().max())
        assert float((g.detach().cpu().double() - w).abs().max()) <= 2e-5 * max(scale, 1.0), (g.shape, scale)


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma"])
@pytest.mark.parametrize("d,k,dil,causal,B,T", [(256, 3, 3, True, 3, 50), (512, 3, 1, False, 2, 41)])
def test_highway_conv_channels_last_matches_channels_first(prec, d, k, dil, causal, B, T):
    """The (B, T, d) layout of ssv_highway_conv_fwd_save / _bwd (what the training graph's highway stacks use between
    layers) gives the output and every gradient of the reference's (B, d, T) layout, bit for bit."""
    from spoofsv_b200 import _lib
    from spoofsv_b200.models.TTSModel import _HighwayConvFn, highwayConv
    torch.manual_seed(3 * d + dil)
    hc = highwayConv(d, k, dil, causal=causal).cuda()
    ps = [hc.conv.weight, hc.conv.bias, hc.ln1.weight, hc.ln1.bias, hc.ln2.weight, hc.ln2.bias]
    p = _lib.PREC_FP32 if prec == "fp32" else _lib.PREC_FP32_FFMA
    x = torch.randn(B, d, T, device="cuda")
    gy = torch.randn(B, d, T, device="cuda")
    outs = []
    for cl in (False, True):
        for q in ps:
            q.grad = None
        xi = (x.transpose(1, 2).contiguous() if cl else x.clone()).requires_grad_(True)
        y = _HighwayConvFn.apply(xi, *ps, k, dil, causal, p, cl)
        y.backward(gy.transpose(1, 2).contiguous() if cl else gy)
        yy, gx = (y.detach().transpose(1, 2), xi.grad.transpose(1, 2)) if cl else (y.detach(), xi.grad)
        outs.append([yy.contiguous(), gx.contiguous()] + [q.grad.clone() for q in ps])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def _grad_close(got, want, tol=2e-5):
    assert got.shape == want.shape
    scale = max(float(want.abs().max()), 1.0)
    assert float((got.detach().cpu().double() - want).abs().max()) <= tol * scale, (tuple(got.shape), scale)


@pytest.mark.parametrize("cin,n,relu_in,with_sb,B,T", [(80, 256, False, True, 3, 37), (256, 256, True, False, 2, 70),
                                                        (512, 256, False, False, 1, 5), (256, 80, True, False, 3, 129),
                                                        (128, 512, False, False, 2, 33), (256, 256, True, True, 4, 9)])
@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma"])
def test_conv_ln_forward_backward_vs_autograd(prec, cin, n, relu_in, with_sb, B, T):
    """ssv_conv_ln_fwd_save / ssv_conv_ln_bwd (the eleven 1x1 conv + LayerNorm layers of Text2Mel,
    models/TTSModel.py:128-131, 173-180, 218-230) against float64 autograd of the same torch statements: output and the
    gradients of the input, the conv, the per-utterance speaker term and the LayerNorm; ragged row counts, the 80-bin
    layers, a batch of one; conv / dgrad / wgrad on the tensor cores (3xTF32, "fp32") and on the CUDA cores."""
    import torch.nn.functional as F
    from spoofsv_b200 import _lib
    from spoofsv_b200.models.TTSModel import _ConvLnFn
    g = torch.Generator().manual_seed(cin + n + T)
    x = torch.randn((B, cin, T), generator=g)
    w = torch.randn((n, cin, 1), generator=g) / cin ** 0.5
    b, gam, bet = (torch.randn((n,), generator=g) * s + o for s, o in ((0.1, 0.0), (0.2, 1.0), (0.1, 0.0)))
    sb = torch.randn((B, n), generator=g) * 0.3 if with_sb else None
    gy = torch.randn((B, n, T), generator=g)
    dev = [t.cuda().requires_grad_(True) for t in (x, w, b, gam, bet)]
    sbd = sb.cuda().requires_grad_(True) if with_sb else None
    y = _ConvLnFn.apply(dev[0], dev[1], dev[2], sbd, dev[3], dev[4], relu_in,
                        _lib.PREC_FP32 if prec == "fp32" else _lib.PREC_FP32_FFMA)
    y.backward(gy.cuda())
    ref = [t.double().requires_grad_(True) for t in (x, w, b, gam, bet)]
    sbr = sb.double().requires_grad_(True) if with_sb else None
    h = F.conv1d(F.relu(ref[0]) if relu_in else ref[0], ref[1], ref[2])
    if with_sb:
        h = h + sbr[:, :, None]
    yo = F.layer_norm(h.transpose(1, 2), (n,), ref[3], ref[4], 1e-5).transpose(1, 2)
    yo.backward(gy.double())
    assert _maxabs(y, yo.float()) <= FP32_TOL
    for got, want in zip(dev, ref):
        _grad_close(got.grad, want.grad)
    if with_sb:
        _grad_close(sbd.grad, sbr.grad)


@pytest.mark.parametrize("B,N,T", [(2, 13, 21), (1, 64, 217), (3, 1, 4)])
def test_train_attention_forward_backward_vs_autograd(B, N, T):
    """ssv_attention_train_fwd / _bwd (models/TTSModel.py:268-272: A = softmax(K^T q / sqrt(d)) over the characters,
    R = V A, [R ; q]) against float64 autograd, with a loss on the alignment as well (the guided-attention term)."""
    from spoofsv_b200.models.TTSModel import _AttentionTrainFn
    g = torch.Generator().manual_seed(B * 100 + N)
    kv = torch.randn((B, 512, N), generator=g)
    q = torch.randn((B, 256, T), generator=g)
    gA, gr = torch.randn((B, N, T), generator=g), torch.randn((B, 512, T), generator=g)
    kvd, qd = kv.cuda().requires_grad_(True), q.cuda().requires_grad_(True)
    A, rq = _AttentionTrainFn.apply(kvd, qd)
    ((A * gA.cuda()).sum() + (rq * gr.cuda()).sum()).backward()
    kvr, qr = kv.double().requires_grad_(True), q.double().requires_grad_(True)
    Ao = torch.softmax(torch.matmul(kvr[:, :256].transpose(1, 2), qr) / 16.0, dim=1)
    ro = torch.cat((torch.matmul(kvr[:, 256:], Ao), qr), dim=1)
    ((Ao * gA.double()).sum() + (ro * gr.double()).sum()).backward()
    assert _maxabs(A, Ao.float()) <= 1e-5 and _maxabs(rq, ro.float()) <= FP32_TOL
    _grad_close(kvd.grad, kvr.grad)
    _grad_close(qd.grad, qr.grad)


def test_text_embedding_and_speaker_projection_vs_autograd():
    """textEmbedding as a gather (models/TTSModel.py:25-35) and the speaker projections (:172-173), forward and
    parameter gradients against float64 autograd of the reference's statements."""
    import torch.nn.functional as F
    from spoofsv_b200.models.TTSModel import _LinearSmallFn, _TextEmbeddingFn
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 34, (3, 17), generator=g)
    W_, b_ = torch.randn((128, 34), generator=g), torch.randn((128,), generator=g)
    gy = torch.randn((3, 128, 17), generator=g)
    wd, bd = W_.cuda().requires_grad_(True), b_.cuda().requires_grad_(True)
    y = _TextEmbeddingFn.apply(ids.cuda(), wd, bd)
    y.backward(gy.cuda())
    wr, br = W_.double().requires_grad_(True), b_.double().requires_grad_(True)
    onehot = F.one_hot(ids, 34).double()                                   # the reference's scatter + Linear
    yo = F.linear(onehot, wr, br).transpose(1, 2)
    yo.backward(gy.double())
    assert _maxabs(y, yo.float()) <= 1e-6
    _grad_close(wd.grad, wr.grad)
    _grad_close(bd.grad, br.grad)

    x = torch.randn((5, 200), generator=g)
    W2, b2 = torch.randn((256, 200), generator=g) / 14, torch.randn((256,), generator=g)
    g2 = torch.randn((5, 256), generator=g)
    w2d, b2d = W2.cuda().requires_grad_(True), b2.cuda().requires_grad_(True)
    s = _LinearSmallFn.apply(x.cuda(), w2d, b2d)
    s.backward(g2.cuda())
    w2r, b2r = W2.double().requires_grad_(True), b2.double().requires_grad_(True)
    so = F.linear(x.double(), w2r, b2r)
    so.backward(g2.double())
    assert _maxabs(s, so.float()) <= 1e-5
    _grad_close(w2d.grad, w2r.grad)
    _grad_close(b2d.grad, b2r.grad)


@pytest.mark.parametrize("B,N,T", [(2, 11, 14), (5, 23, 131)])
def test_text2mel_training_backward_vs_autograd(B, N, T, cuda_models_k):
    """loss.backward() through the train branch (train/adversarial_wasserstein_gp.py:277-300): every parameter
    gradient of Text2Mel against float64 autograd through the oracle's restatement, on the same weights; a tiny
    shape and one longer than every receptive field, with several row tiles and wgrad chunks."""
    m1, _, sd1, _ = cuda_models_k
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(B, N, seed=5)
    spk = torch.from_numpy(emb[:B].copy())[:, :, None]
    mel = torch.rand((B, 80, T), generator=torch.Generator().manual_seed(4))
    tgt = torch.rand((B, 80, T), generator=torch.Generator().manual_seed(6))
    m1.train()
    try:
        m1.zero_grad(set_to_none=True)
        Y, A = m1(mel.cuda(), ids.cuda(), spk.cuda())
        assert Y.requires_grad
        loss = (Y - tgt.cuda()).abs().mean() + 0.1 * (A * A).mean()
        loss.backward()
        got = {n: p.grad.detach().cpu().double() for n, p in m1.named_parameters()}
    finally:
        m1.eval()
        m1.zero_grad(set_to_none=True)
    sd = {k: v.double().requires_grad_(True) for k, v in sd1.items()}
    oY, oA = O.melsyn_train_forward(sd, mel.double(), ids, spk.double())
    assert _maxabs(Y, oY.float()) <= FP32_TOL and _maxabs(A, oA.float()) <= FP32_TOL
    ((oY - tgt.double()).abs().mean() + 0.1 * (oA * oA).mean()).backward()
    assert set(got) == set(sd)
    worst = 0.0
    for n, g in got.items():
        w = sd[n].grad
        assert w is not None and g.shape == w.shape, n
        err = float((g - w).abs().max()) / max(float(w.abs().max()), 1e-6)
        worst = max(worst, err)
        assert err <= 2e-3, (n, err)
    print(f"training backward: worst relative gradient error {worst:.2e}")


def test_adversarial_training_step_golden(golden_dir, cuda_models_k):
    """One generator and one discriminator iteration of train/adversarial_wasserstein_gp.py:269-316 against the
    unmodified reference (tests/golden/train_step_seed7.npz, oracle/make_golden.py --train-step-only): loss terms,
    generator gradients (highway convs in the library's kernels), gradient penalty, discriminator gradients."""
    from spoofsv_b200 import train as TR
    g = np.load(golden_dir / "train_step_seed7.npz")
    m1 = cuda_models_k[0]
    disc = TR.melDisc(80, 128).cuda()
    disc.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("disc/")})
    disc.eval()                                        # dropout off, as in the fixture
    mel_gt, ids, spk = (torch.from_numpy(g[k]).cuda() for k in ("mel_gt", "textid", "spk"))
    cfg = {"LAMBDA": 10}
    gaw = TR.guided_attention_mat(186, 325, device="cuda")
    opt_g = torch.optim.SGD(m1.parameters(), lr=0.0)   # the step itself leaves the shared fixture model untouched
    opt_d = torch.optim.SGD(disc.parameters(), lr=0.0)
    m1.train()
    m1.precision = "fp32-ffma"          # the CUDA-core arm: 1e-4 of the reference's own fp32 run on every gradient
    try:
        terms = TR.generator_step(m1, disc, opt_g, mel_gt, ids, spk, gaw, cfg)
        want = g["g_terms"]
        for k, w in zip(("l1", "bin_div", "att", "disc"), want[:4]):
            assert abs(terms[k] - w) <= 1e-5 * max(1.0, abs(w)), (k, terms[k], w)
        grads = dict(m1.named_parameters())
        for k in g.files:
            if not k.startswith("ggrad/"):
                continue
            got = grads[k[6:]].grad.detach().cpu().numpy()
            ref = g[k]
            if got.size != ref.size:
                got = got.reshape(-1)[::37]
            err = np.abs(got.reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-6)
            print(f"G grad {k[6:]}: rel err {err:.2e}")
            assert err <= 1e-4, k                   # the fixture is the reference's own fp32 CPU run
        dterms = TR.discriminator_step(m1, disc, opt_d, mel_gt, ids, spk, cfg, coeff=torch.from_numpy(g["coeff"]))
        assert abs(dterms["gp"] - g["d_terms"][0]) <= 1e-4 * abs(g["d_terms"][0])
        assert abs(dterms["wd"] - g["d_terms"][1]) <= 1e-5
        dgr = dict(disc.named_parameters())
        for k in g.files:
            if not k.startswith("dgrad/"):
                continue
            got = dgr[k[6:]].grad.detach().cpu().numpy()
            ref = g[k]
            if got.size != ref.size:
                got = got.reshape(-1)[::37]
            err = np.abs(got.reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-6)
            print(f"D grad {k[6:]}: rel err {err:.2e}")
            assert err <= 1e-4, k
        # the tensor-core arm (highway convs' forward and dgrad as 3xTF32 tcgen05 MMAs): the same loss terms; gradients to
        # 2e-3 of their largest element (a ReLU gate at |x| ~ 1e-6 flips here or there: the same effect separates torch's
        # own fp32 GPU run from its CPU run on this model)
        m1.precision = "fp32"
        m1.zero_grad(set_to_none=True)
        terms = TR.generator_step(m1, disc, opt_g, mel_gt, ids, spk, gaw, cfg)
        for k, w in zip(("l1", "bin_div", "att", "disc"), g["g_terms"][:4]):
            assert abs(terms[k] - w) <= 2e-5 * max(1.0, abs(w)), (k, terms[k], w)
        worst = 0.0
        for k in g.files:
            if not k.startswith("ggrad/"):
                continue
            got = grads[k[6:]].grad.detach().cpu().numpy()
            ref = g[k]
            if got.size != ref.size:
                got = got.reshape(-1)[::37]
            err = np.abs(got.reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-6)
            worst = max(worst, err)
            assert err <= 2e-3, (k, err)
        print(f"tensor-core arm: worst relative gradient error {worst:.2e}")
    finally:
        m1.precision = "fp32"
        m1.eval()
        m1.zero_grad(set_to_none=True)


def test_ssrn_training_backward_vs_autograd(cuda_models_k):
    """SSRN in train() mode (the `train_ssrn` step, train/adversarial_wasserstein_gp.py:324-338): forward and
    every parameter gradient against float64 autograd through the oracle's restatement of SSRN.forward."""
    _, m2, _, sd2 = cuda_models_k
    B, T = 2, 19
    mel = torch.rand((B, 80, T), generator=torch.Generator().manual_seed(8))
    tgt = torch.rand((B, 513, 4 * T), generator=torch.Generator().manual_seed(9))
    prec = m2.precision
    m2.precision = "fp32"
    m2.train()
    try:
        m2.zero_grad(set_to_none=True)
        Y = m2(mel.cuda())
        assert Y.requires_grad and tuple(Y.shape) == (B, 513, 4 * T)
        t = tgt.cuda()
        ((Y - t).abs().mean() + torch.mean(-t * torch.log(Y + 1e-8) - (1 - t) * torch.log(1 - Y + 1e-8))).backward()
        got = {n: p.grad.detach().cpu().double() for n, p in m2.named_parameters()}
    finally:
        m2.eval()
        m2.precision = prec
        m2.zero_grad(set_to_none=True)
    sd = {k: v.double().requires_grad_(True) for k, v in sd2.items()}
    oY = O.ssrn(mel.double(), sd)
    assert _maxabs(Y, oY.float()) <= FP32_TOL
    td = tgt.double()
    ((oY - td).abs().mean() + torch.mean(-td * torch.log(oY + 1e-8) - (1 - td) * torch.log(1 - oY + 1e-8))).backward()
    assert set(got) == set(sd)
    worst = 0.0
    for n, g in got.items():
        w = sd[n].grad
        err = float((g - w).abs().max()) / max(float(w.abs().max()), 1e-9)
        worst = max(worst, err)
        assert err <= 1e-3, (n, err)
    print(f"SSRN training backward: worst relative gradient error {worst:.2e}")


def test_ssrn_adversarial_step_vs_oracle(cuda_models_k):
    """`train_ssrn` iterations (train/adversarial_wasserstein_gp.py:324-360): loss terms of the generator step and
    gradient penalty / Wasserstein distance of the discriminator step against the same statements evaluated on the
    oracle's SSRN in float64 on the CPU."""
    from spoofsv_b200 import train as TR
    _, m2, _, sd2 = cuda_models_k
    B, T = 2, 19
    g = torch.Generator().manual_seed(12)
    mel = torch.rand((B, 80, T), generator=g)
    lin = torch.rand((B, 513, 4 * T), generator=g) * 0.9 + 0.05
    coeff = torch.rand(B, generator=g)
    torch.manual_seed(3)
    disc = TR.linDisc(513, 128).eval()
    ref_disc = TR.linDisc(513, 128).double().eval()
    ref_disc.load_state_dict({k: v.double() for k, v in disc.state_dict().items()})
    disc = disc.cuda()

    class _NoStep:
        def __init__(self, ps): self.ps = list(ps)
        def zero_grad(self, set_to_none=True):
            for p in self.ps: p.grad = None
        def step(self): pass

    prec = m2.precision
    m2.precision = "fp32"
    m2.train()
    try:
        gt = TR.ssrn_generator_step(m2, disc, _NoStep(m2.parameters()), mel.cuda(), lin.cuda())
        dt = TR.ssrn_discriminator_step(m2, disc, _NoStep(disc.parameters()), mel.cuda(), lin.cuda(), {"LAMBDA": 10}, coeff=coeff)
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in disc.parameters())
    finally:
        m2.eval()
        m2.precision = prec
        m2.zero_grad(set_to_none=True)
    with torch.no_grad():
        pred = O.ssrn(mel.double(), {k: v.double() for k, v in sd2.items()})
    l1 = float(torch.mean(torch.abs(lin.double() - pred)))
    bd = float(torch.mean(-lin.double() * torch.log(pred + 1e-8) - (1 - lin.double()) * torch.log(1 - pred + 1e-8)))
    with torch.no_grad():
        ld = float(torch.mean(-ref_disc(pred)))
    assert abs(gt["l1"] - l1) <= 1e-5 and abs(gt["bin_div"] - bd) <= 1e-5 and abs(gt["disc"] - ld) <= 1e-4 * max(1.0, abs(ld))
    c = coeff.double()[:, None, None]
    mid = (c * lin.double() + (1 - c) * pred).requires_grad_(True)
    gr = torch.autograd.grad(ref_disc(mid).sum(), mid)[0]
    gp = float(torch.mean(10 * (torch.norm(gr, p=2, dim=(1, 2)) - 1) ** 2))
    with torch.no_grad():
        wd = -float(torch.mean(ref_disc(pred) - ref_disc(lin.double())))
    assert abs(dt["gp"] - gp) <= 1e-4 * max(1.0, abs(gp)) and abs(dt["wd"] - wd) <= 1e-4 * max(1.0, abs(wd))


def test_graphed_training_iterations_match_eager(cuda_models_k):
    """train.GraphedIteration: the G and the D iteration captured in CUDA graphs and replayed give the loss terms and the
    updated parameters of the eager iterations from the same state (dropout off: the discriminator must be
    deterministic for the comparison; Adam built capturable)."""
    from spoofsv_b200 import train as TR
    from oracle import weights as Wt
    m1, _, _, _ = cuda_models_k
    sd0 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    B, N, T = 4, 24, 48
    gen = torch.Generator(device="cuda").manual_seed(5)
    ids = Wt.synthetic_text(B, N, seed=9).cuda()
    spk = torch.randn((B, m1.spkemb_dim, 1), device="cuda", generator=gen) * 0.1
    mel = torch.rand((B, 80, T), device="cuda", generator=gen) * 0.9 + 0.05
    coeff = torch.rand(B, device="cuda", generator=gen)
    torch.manual_seed(11)
    disc = TR.melDisc(80, 128).cuda().train()
    for mod in disc.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    gaw = TR.guided_attention_mat(186, 325, device="cuda")
    cfg = {"LAMBDA": 10}
    og = torch.optim.Adam(m1.parameters(), 2e-4, (0.5, 0.9), 1e-6, capturable=True)
    od = torch.optim.Adam(disc.parameters(), 2e-4, (0.5, 0.9), 1e-6, capturable=True)

    def snapshot(mod, opt):
        out = [p.detach().clone() for p in mod.parameters()]
        for st in opt.state.values():
            out += [v.clone() for v in st.values() if torch.is_tensor(v)]
        return out

    def restore(mod, opt, snap):
        it = iter(snap)
        with torch.no_grad():
            for p in mod.parameters():
                p.copy_(next(it))
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.copy_(next(it))

    was_training = m1.training
    try:
        m1.train()
        m1.layerwise_train_forward = True
        for body, mod, opt, inputs, to_dict in (
                (lambda me, i, sp: TR.generator_body(m1, disc, og, me, i, sp, gaw, cfg), m1, og, (mel, ids, spk), TR.generator_terms),
                (lambda me, i, sp, c: TR.discriminator_body(m1, disc, od, me, i, sp, cfg, coeff=c), disc, od, (mel, ids, spk, coeff),
                 TR.wgan_terms)):
            g = TR.GraphedIteration(body, inputs, parameters=list(m1.parameters()) + list(disc.parameters()))
            snap = snapshot(mod, opt)
            got = to_dict(g(*inputs))
            p_graph = [p.detach().clone() for p in mod.parameters()]
            restore(mod, opt, snap)
            want = to_dict(body(*inputs))
            for k in want:
                assert abs(got[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])
            for a, b in zip(p_graph, mod.parameters()):
                assert float((a - b.detach()).abs().max()) <= 1e-7
            del g
    finally:
        m1.layerwise_train_forward = False
        m1.train(was_training)
        m1.zero_grad(set_to_none=True)
        m1.load_state_dict(sd0)
