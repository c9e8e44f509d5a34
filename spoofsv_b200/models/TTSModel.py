"""Drop-in ``melSyn`` / ``SSRN`` / ``highwayConv`` backed by the sm_100a CUDA library.

Same constructor arguments, forward signatures, return tuples and ``state_dict``
key layout as the reference ``models/TTSModel.py`` (melSyn :234-300, SSRN :319-362,
highwayConv :37-84), so ``*.tar.pth`` checkpoints load with
``load_state_dict(ckp['model_state_dict'])`` unchanged.  The ``torch.nn`` submodules
below only *hold parameters* (and reproduce the reference's construction order, so
``torch.manual_seed(s)`` yields identical random-init weights); no torch math runs in
``forward`` -- it hands raw device pointers to the C ABI (include/spoofsv_b200.h).

There is no CPU fallback: a CPU tensor, a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib

__all__ = ["melSyn", "SSRN", "highwayConv"]


class _Bag(nn.Module):
    """Ordered parameter container; children are registered in keyword order."""

    def __init__(self, **children: nn.Module):
        super().__init__()
        for name, child in children.items():
            self.add_module(name, child)

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("parameter container; call the owning model instead")


def _hc(d: int, k: int) -> _Bag:
    return _Bag(conv=nn.Conv1d(d, 2 * d, k), ln1=nn.LayerNorm(d), ln2=nn.LayerNorm(d))


def _hci(d: int) -> _Bag:
    return _Bag(hc1=_hc(d, 3), hc2=_hc(d, 3), hc3=_hc(d, 3), hc4=_hc(d, 3))


def _prec(name: str) -> int:
    try:
        return {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp32-ffma": _lib.PREC_FP32_FFMA}[name]
    except KeyError:
        raise ValueError(f"precision must be 'fp32', 'fp32-ffma' or 'bf16', got {name!r}") from None


class _Native(nn.Module):
    """Owns a native handle built from the module's parameters; rebuilt when they change."""

    _create_name = ""
    _destroy_name = ""

    def __init__(self):
        super().__init__()
        self._handle: Optional[int] = None
        self._stamp = None

    def _dims(self):  # pragma: no cover
        raise NotImplementedError

    def _param_stamp(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def _release(self):
        if self._handle is not None:
            try:
                getattr(_lib.load(False), self._destroy_name)(C.c_void_p(self._handle))
            except Exception:
                pass
            self._handle = None

    def _native(self) -> C.c_void_p:
        stamp = self._param_stamp()
        if self._handle is None or stamp != self._stamp:
            self._release()
            self._on_rebuild()
            lib = _lib.load()
            names, ptrs, numels, n, keep = _lib.pack_params(self.state_dict())
            out = C.c_void_p()
            _lib.check(getattr(lib, self._create_name)(names, ptrs, numels, n, *self._dims(), C.byref(out)))
            del keep
            self._handle = out.value
            self._stamp = stamp
        return C.c_void_p(self._handle)

    def _on_rebuild(self):
        pass

    def __del__(self):
        try:
            self._release()
        except Exception:   # interpreter shutdown: torch internals may already be gone
            pass


class highwayConv(nn.Module):
    """reference models/TTSModel.py:37-84 -- same ctor, same parameters, CUDA forward."""

    def __init__(self, dimension, kernel_size, dilation, causal=False):
        super().__init__()
        self.dimension, self.kernel_size, self.dilation, self.causal = dimension, kernel_size, dilation, causal
        self.pad = dilation * (kernel_size - 1) // 2
        self.conv = nn.Conv1d(dimension, 2 * dimension, kernel_size, padding=0 if causal else self.pad,
                              dilation=dilation)
        self.ln1 = nn.LayerNorm(dimension)
        self.ln2 = nn.LayerNorm(dimension)
        self.precision = "fp32"

    def forward(self, inputs):
        _lib.require_cuda(inputs, "highwayConv.forward")
        if inputs.dim() != 3 or inputs.shape[1] != self.dimension:
            raise ValueError(f"highwayConv expects (B, {self.dimension}, T), got {tuple(inputs.shape)}")
        ps = [self.conv.weight, self.conv.bias, self.ln1.weight, self.ln1.bias, self.ln2.weight, self.ln2.bias]
        for p in ps:
            _lib.require_cuda(p, "highwayConv parameter")
        # the backward pass exists for the FP32 arm; the bf16 tensor-core arm is inference-only (no autograd graph)
        if self.precision in ("fp32", "fp32-ffma") and torch.is_grad_enabled() and (inputs.requires_grad or any(p.requires_grad for p in ps)):
            return _HighwayConvFn.apply(inputs, *ps, self.kernel_size, self.dilation, bool(self.causal), _prec(self.precision))
        return _highway_fwd(inputs, ps, self.kernel_size, self.dilation, self.causal, self.precision)


def _highway_fwd(inputs, ps, k, dilation, causal, precision):
    x = inputs.detach().to(torch.float32).contiguous()
    B, d, T = x.shape
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    ps = [p.detach().contiguous() for p in ps]
    _lib.check(_lib.load().ssv_highway_conv_fwd(
        x.data_ptr(), *[p.data_ptr() for p in ps], B, d, T, k, dilation, int(causal),
        y.data_ptr(), _prec(precision), _lib.current_stream_ptr()))
    return y


class _HighwayConvFn(torch.autograd.Function):
    """highwayConv with a hand-written backward (ssv_highway_conv_bwd): what autograd derives for
    models/TTSModel.py:63-84 -- gate, both LayerNorms, dgrad and wgrad of the dilated conv -- in FP32.  prec: the conv and
    its dgrad on the tensor cores with split operands (PREC_FP32, FP32-accurate) or on the CUDA cores (PREC_FP32_FFMA)."""

    save_h = True        # keep H = conv(x) + b for the backward pass (2d floats per row) instead of recomputing it

    @staticmethod
    def forward(ctx, x, w, b, g1, b1, g2, b2, k, dilation, causal, prec=0, channels_last=False):
        """channels_last: x and the result are (B, T, d) -- the library's row layout -- so a stack of layers hands its
        activations on without the boundary transposes."""
        ps = [p.detach().contiguous() for p in (w, b, g1, b1, g2, b2)]
        xs = x.detach().to(torch.float32).contiguous()
        if channels_last:
            B, T, d = xs.shape
        else:
            B, d, T = xs.shape
        h = None
        if xs.numel() == 0:
            y = torch.empty_like(xs)
        elif _HighwayConvFn.save_h:
            y = torch.empty_like(xs)
            h = torch.empty((B * T, 2 * d), device=xs.device, dtype=torch.float32)
            _lib.check(_lib.load().ssv_highway_conv_fwd_save(
                xs.data_ptr(), *[p.data_ptr() for p in ps], B, d, T, k, dilation, int(causal), y.data_ptr(), h.data_ptr(),
                int(prec), int(channels_last), _lib.current_stream_ptr()))
        else:
            xin = xs.transpose(1, 2).contiguous() if channels_last else xs
            y = _highway_fwd(xin, ps, k, dilation, causal, "fp32-ffma" if prec == _lib.PREC_FP32_FFMA else "fp32")
            if channels_last:
                y = y.transpose(1, 2).contiguous()
        ctx.save_for_backward(xs, *ps, *([h] if h is not None else []))
        ctx.cfg = (k, dilation, causal, h is not None, int(prec), bool(channels_last), (B, d, T))
        return y

    @staticmethod
    def backward(ctx, dy):
        k, dilation, causal, has_h, prec, channels_last, (B, d, T) = ctx.cfg
        saved = ctx.saved_tensors
        x, w, b, g1, b1, g2, b2 = saved[:7]
        h = saved[7] if has_h else None
        dy = dy.detach().to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        grads = [torch.empty_like(p) for p in (w, b, g1, b1, g2, b2)]
        if x.numel() == 0:
            return (dx, *[torch.zeros_like(p) for p in (w, b, g1, b1, g2, b2)], None, None, None, None, None)
        _lib.check(_lib.load().ssv_highway_conv_bwd(
            x.data_ptr(), dy.data_ptr(), w.data_ptr(), b.data_ptr(), g1.data_ptr(), b1.data_ptr(), g2.data_ptr(),
            b2.data_ptr(), B, d, T, k, dilation, int(causal), h.data_ptr() if h is not None else None, dx.data_ptr(),
            *[g.data_ptr() for g in grads], prec, int(channels_last), _lib.current_stream_ptr()))
        return (dx, *grads, None, None, None, None, None)


class _ConvLnFn(torch.autograd.Function):
    """1x1 conv (+ per-utterance bias) + LayerNorm with a hand-written backward (ssv_conv_ln_fwd_save / ssv_conv_ln_bwd):
    the eleven plain layers of Text2Mel (models/TTSModel.py:128-131, 173-180, 218-230), FP32.  relu_in folds the ReLU the
    reference applies to the layer's input."""

    @staticmethod
    def forward(ctx, x, w, b, sb, g, beta, relu_in, prec=_lib.PREC_FP32_FFMA):
        xs = x.detach().to(torch.float32).contiguous()
        ws, bs, gs, betas = (p.detach().contiguous() for p in (w, b, g, beta))
        sbs = None if sb is None else sb.detach().to(torch.float32).contiguous()
        B, cin, T = xs.shape
        n = ws.shape[0]
        y = torch.empty((B, n, T), device=xs.device, dtype=torch.float32)
        h = torch.empty((B * T, (n + 63) // 64 * 64), device=xs.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_conv_ln_fwd_save(
            xs.data_ptr(), ws.data_ptr(), bs.data_ptr(), None if sbs is None else sbs.data_ptr(), gs.data_ptr(), betas.data_ptr(),
            B, cin, n, T, int(relu_in), y.data_ptr(), h.data_ptr(), int(prec), _lib.current_stream_ptr()))
        ctx.save_for_backward(xs, ws, gs, h)
        ctx.cfg = (bool(relu_in), sbs is not None, ctx.needs_input_grad[0], int(prec))
        return y

    @staticmethod
    def backward(ctx, dy):
        relu_in, has_sb, need_dx, prec = ctx.cfg
        x, w, g, h = ctx.saved_tensors
        B, cin, T = x.shape
        n = w.shape[0]
        dy = dy.detach().to(torch.float32).contiguous()
        dx = torch.empty_like(x) if need_dx else None
        dw = torch.empty_like(w)
        db = torch.empty((n,), device=x.device, dtype=torch.float32)
        dsb = torch.empty((B, n), device=x.device, dtype=torch.float32) if has_sb else None
        dg, dbeta = torch.empty_like(db), torch.empty_like(db)
        _lib.check(_lib.load().ssv_conv_ln_bwd(
            x.data_ptr(), dy.data_ptr(), w.data_ptr(), g.data_ptr(), h.data_ptr(), B, cin, n, T, int(relu_in),
            None if dx is None else dx.data_ptr(), dw.data_ptr(), db.data_ptr(), None if dsb is None else dsb.data_ptr(),
            dg.data_ptr(), dbeta.data_ptr(), prec, _lib.current_stream_ptr()))
        return dx, dw, db, dsb, dg, dbeta, None, None


class _AttentionTrainFn(torch.autograd.Function):
    """Unmasked softmax attention of the train branch (models/TTSModel.py:268-272): kv = [K ; V] (B, 512, N),
    q (B, 256, T) -> A (B, N, T), [V A ; q] (B, 512, T); backward in ssv_attention_train_bwd."""

    @staticmethod
    def forward(ctx, kv, q):
        kvs, qs = kv.detach().to(torch.float32).contiguous(), q.detach().to(torch.float32).contiguous()
        B, _, N = kvs.shape
        T = qs.shape[-1]
        A = torch.empty((B, N, T), device=qs.device, dtype=torch.float32)
        rq = torch.empty((B, 512, T), device=qs.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_attention_train_fwd(kvs.data_ptr(), qs.data_ptr(), B, N, T, A.data_ptr(), rq.data_ptr(),
                                                       _lib.current_stream_ptr()))
        ctx.save_for_backward(kvs, qs, A)
        return A, rq

    @staticmethod
    def backward(ctx, dA, drq):
        kv, q, A = ctx.saved_tensors
        B, _, N = kv.shape
        T = q.shape[-1]
        drq = (torch.zeros((B, 512, T), device=q.device) if drq is None else drq.detach().to(torch.float32)).contiguous()
        dA = None if dA is None else dA.detach().to(torch.float32).contiguous()
        dkv, dq = torch.empty_like(kv), torch.empty_like(q)
        _lib.check(_lib.load().ssv_attention_train_bwd(kv.data_ptr(), q.data_ptr(), A.data_ptr(), None if dA is None else dA.data_ptr(),
                                                       drq.data_ptr(), B, N, T, dkv.data_ptr(), dq.data_ptr(),
                                                       _lib.current_stream_ptr()))
        return dkv, dq


class _TextEmbeddingFn(torch.autograd.Function):
    """textEmbedding (models/TTSModel.py:25-35) as a column gather of W.weight + bias; backward sums per vocabulary id."""

    @staticmethod
    def forward(ctx, ids, weight, bias):
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        idc = ids.detach().to(torch.int64).contiguous()
        B, N = idc.shape
        E, vocab = w.shape
        y = torch.empty((B, E, N), device=w.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_text_embedding_fwd(idc.data_ptr(), w.data_ptr(), b.data_ptr(), B, N, vocab, E, y.data_ptr(),
                                                      _lib.current_stream_ptr()))
        ctx.save_for_backward(idc)
        ctx.shape = (E, vocab)
        return y

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        E, vocab = ctx.shape
        B, N = ids.shape
        dy = dy.detach().to(torch.float32).contiguous()
        dw = torch.empty((E, vocab), device=dy.device, dtype=torch.float32)
        db = torch.empty((E,), device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_text_embedding_bwd(ids.data_ptr(), dy.data_ptr(), B, N, vocab, E, dw.data_ptr(), db.data_ptr(),
                                                      _lib.current_stream_ptr()))
        return None, dw, db


class _LinearSmallFn(torch.autograd.Function):
    """Speaker projection fc1 / fc2 (models/TTSModel.py:172-173): (B, in) -> (B, out); the input needs no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xs, w, b = x.detach().to(torch.float32).contiguous(), weight.detach().contiguous(), bias.detach().contiguous()
        B, in_f = xs.shape
        out_f = w.shape[0]
        y = torch.empty((B, out_f), device=xs.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_linear_small_fwd(xs.data_ptr(), w.data_ptr(), b.data_ptr(), B, in_f, out_f, y.data_ptr(),
                                                    _lib.current_stream_ptr()))
        ctx.save_for_backward(xs)
        ctx.shape = (in_f, out_f)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        in_f, out_f = ctx.shape
        dy = dy.detach().to(torch.float32).contiguous()
        dw = torch.empty((out_f, in_f), device=x.device, dtype=torch.float32)
        db = torch.empty((out_f,), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().ssv_linear_small_bwd(x.data_ptr(), dy.data_ptr(), x.shape[0], in_f, out_f, dw.data_ptr(), db.data_ptr(),
                                                    _lib.current_stream_ptr()))
        return None, dw, db


class melSyn(_Native):
    """Text2Mel.  reference models/TTSModel.py:234-300."""

    _create_name = "ssv_text2mel_create"
    _destroy_name = "ssv_text2mel_destroy"

    def __init__(self, vocab_len, condition, spkemb_dim, textemb_dim=128, freq_bins=80, hidden_dim=256):
        super().__init__()
        if not condition:
            raise NotImplementedError("spoofsv_b200 implements the speaker-conditioned model (condition=True)")
        h, h2 = hidden_dim, 2 * hidden_dim
        self.vocab_len, self.spkemb_dim, self.textemb_dim = vocab_len, spkemb_dim, textemb_dim
        self.freq_bins, self.hidden_dim = freq_bins, hidden_dim
        pw = lambda i, o: nn.Conv1d(i, o, 1)
        # construction order == reference order (fixes the RNG stream and the state_dict key order)
        self.text_encoder = _Bag(
            textemb_layer=_Bag(W=nn.Linear(vocab_len, textemb_dim)),
            conv1=pw(textemb_dim, h2), ln1=nn.LayerNorm(h2), conv2=pw(h2, h2), ln2=nn.LayerNorm(h2),
            hci1=_hci(h2), hci2=_hci(h2), hc1=_hc(h2, 3), hc2=_hc(h2, 3), hc3=_hc(h2, 1), hc4=_hc(h2, 1))
        self.audio_encoder = _Bag(
            fc1=nn.Linear(spkemb_dim, h), fc2=nn.Linear(spkemb_dim, h),
            conv1=pw(freq_bins, h), ln1=nn.LayerNorm(h), conv2=pw(h, h), ln2=nn.LayerNorm(h),
            conv3=pw(h, h), ln3=nn.LayerNorm(h),
            hci1=_hci(h), hci2=_hci(h), hc1=_hc(h, 3), hc2=_hc(h, 3))
        self.audio_decoder = _Bag(
            conv1=pw(h2, h), ln1=nn.LayerNorm(h), hci=_hci(h), hc1=_hc(h, 3), hc2=_hc(h, 3),
            conv2=pw(h, h), ln2=nn.LayerNorm(h), conv3=pw(h, h), ln3=nn.LayerNorm(h),
            conv4=pw(h, h), ln4=nn.LayerNorm(h), conv5=pw(h, freq_bins), ln5=nn.LayerNorm(freq_bins))
        self.precision = "fp32"
        self.max_frames = 1024      # decoder capacity (reference MAX_FRAME_NUM is 325)
        self.decode_plan = None     # (rows per micro-batch, warps per row[, front-end warps per CTA]) override of the
                                    # decode kernel's front-end shape, None / 0 = measured choice
                                    # (ssv_decoder_set_plan; tuning and tests)
        self._dec: Optional[int] = None
        self._dec_cap = (0, 0, 0)
        self._state = None

    # ---- native handles ------------------------------------------------------------------
    def _dims(self):
        return (self.vocab_len, self.spkemb_dim, self.textemb_dim, self.freq_bins, self.hidden_dim)

    def _drop_decoder(self):
        if self._dec is not None:
            try:
                _lib.load(False).ssv_decoder_destroy(C.c_void_p(self._dec))
            except Exception:
                pass
            self._dec = None
            self._dec_cap = (0, 0, 0)
        self._state = None

    def _on_rebuild(self):
        self._drop_decoder()

    def _release(self):
        self._drop_decoder()
        super()._release()

    def _decoder(self, B: int, N: int, T: int) -> C.c_void_p:
        h = self._native()
        cb, cn, ct = self._dec_cap
        if self._dec is None or B > cb or N > cn or T > ct:
            self._drop_decoder()
            cap = (max(B, cb), max(N, cn), max(T, ct))
            out = C.c_void_p()
            _lib.check(_lib.load().ssv_decoder_create(h, *cap, C.byref(out)))
            self._dec, self._dec_cap = out.value, cap
        r, w, f = (tuple(self.decode_plan or ()) + (0, 0, 0))[:3]
        _lib.check(_lib.load().ssv_decoder_set_plan(C.c_void_p(self._dec), int(r or 0), int(w or 0), int(f or 0)))
        return C.c_void_p(self._dec)

    # ---- pieces --------------------------------------------------------------------------
    def encode_text(self, textid: torch.Tensor, check: bool = True):
        """textEncoder.forward (reference :126-140): (B, 1, N) int -> K, V (B, hidden, N).

        The id range is verified on the device by the embedding gather (no host round trip before the launch);
        ``check=True`` synchronises afterwards and raises ValueError for an id outside [0, vocab_len), as the
        reference's one-hot scatter would.  The decode paths pass ``check=False``: ``ssv_decoder_check`` at the end of
        the utterance batch reports the same flag."""
        _lib.require_cuda(textid, "melSyn text ids")
        ids = textid.detach().to(torch.int64).contiguous()
        if ids.dim() != 3 or ids.shape[1] != 1:
            raise ValueError(f"textid must be (B, 1, N), got {tuple(ids.shape)}")
        B, _, N = ids.shape
        if B == 0 or N == 0:
            raise ValueError("textid is empty")
        K = torch.empty((B, self.hidden_dim, N), device=ids.device, dtype=torch.float32)
        V = torch.empty_like(K)
        h = self._native()
        _lib.check(_lib.load().ssv_text_encoder_fwd(h, ids.data_ptr(), B, N, K.data_ptr(), V.data_ptr(),
                                                    _prec(self.precision), _lib.current_stream_ptr()))
        if check:
            _lib.check(_lib.load().ssv_text2mel_check(h, _lib.current_stream_ptr()))
        return K, V

    def _begin(self, K, V, spkemb, t_cap: int):
        B, _, N = K.shape
        spk = spkemb.detach().to(torch.float32).contiguous()
        if spk.shape != (B, self.spkemb_dim, 1):
            raise ValueError(f"spkemb must be ({B}, {self.spkemb_dim}, 1), got {tuple(spk.shape)}")
        dec = self._decoder(B, N, t_cap)
        dev = K.device
        Y = torch.empty((B, self.freq_bins, t_cap), device=dev, dtype=torch.float32)
        A = torch.empty((B, N, t_cap), device=dev, dtype=torch.float32)
        traj = torch.empty((t_cap, B), device=dev, dtype=torch.int64)
        Kc, Vc = K.detach().contiguous(), V.detach().contiguous()
        _lib.check(_lib.load().ssv_decoder_begin(dec, Kc.data_ptr(), Vc.data_ptr(), spk.data_ptr(), B, N,
                                                 Y.data_ptr(), A.data_ptr(), traj.data_ptr(), t_cap,
                                                 _lib.current_stream_ptr()))
        self._state = dict(Y=Y, A=A, traj=traj, B=B, N=N, t=0, cap=t_cap, keep=(Kc, Vc, spk))
        return dec

    def synthesize(self, textid, spkemb, n_frames: int):
        """Whole AR loop of generate_test_utterances.py:105-116 in one launch.

        Returns (Y (B,F,T), A (B,N,T), pma trajectory (T,B), K, V)."""
        if n_frames < 1:
            raise ValueError("n_frames must be >= 1")
        K, V = self.encode_text(textid, check=False)        # bad ids surface in ssv_decoder_check below
        dec = self._begin(K, V, spkemb, n_frames)
        lib = _lib.load()
        _lib.check(lib.ssv_decoder_run(dec, n_frames, _lib.current_stream_ptr()))
        _lib.check(lib.ssv_decoder_check(dec, _lib.current_stream_ptr()))
        st = self._state
        st["t"] = n_frames
        return st["Y"], st["A"], st["traj"], K, V

    # ---- reference protocol --------------------------------------------------------------
    def forward(self, melspec, textid, spkemb, K=None, V=None, A_last=None, pma=None):
        """Eval-mode protocol of the reference (:275-300).

        Call 1: melspec (B,F,1) -> (Y, A, max_att, K, V).  Call t: melspec (B,F,t) whose newest
        column is the next input frame -> (Y (B,F,t), A (B,N,t), max_att).  Y and A are views of
        decoder-owned buffers.  ``A_last`` is not read (its columns are already held here)."""
        if self.training:
            # layerwise_train_forward: also without a graph (the discriminator iteration) go layer by layer instead of
            # through the fused forward, whose packed copy of the weights is rebuilt on the host after every optimizer
            # step -- a CUDA-graph capture of the iteration (train.GraphedIteration) cannot contain that
            layerwise = getattr(self, "layerwise_train_forward", False) or (
                torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))
            if self.precision in ("fp32", "fp32-ffma") and layerwise:
                return self._train_forward_autograd(melspec, textid, spkemb)
            return self._train_forward(melspec, textid, spkemb)
        _lib.require_cuda(melspec, "melSyn.forward melspec")
        if melspec.dim() != 3 or melspec.shape[1] != self.freq_bins:
            raise ValueError(f"melspec must be (B, {self.freq_bins}, T), got {tuple(melspec.shape)}")
        B, _, T = melspec.shape
        if T < 1 or B < 1:
            raise ValueError("melspec is empty")
        lib = _lib.load()
        first = T == 1
        if first:
            if textid is None:
                raise ValueError("the first call (T == 1) needs textid")
            K, V = self.encode_text(textid)
            if K.shape[0] != B:
                raise ValueError("batch size of textid and melspec differ")
            dec = self._begin(K, V, spkemb, self.max_frames)
        else:
            st = self._state
            if st is None or st["B"] != B or st["t"] != T - 1:
                have = None if st is None else st["t"]
                raise RuntimeError(
                    f"melSyn.forward: incremental decoder holds {have} frame(s) but was called with T={T}; "
                    "the reference protocol appends exactly one frame per call after a T == 1 call")
            if T > st["cap"]:
                raise RuntimeError(f"melSyn.forward: T={T} exceeds max_frames={st['cap']}; raise model.max_frames")
            dec = C.c_void_p(self._dec)
        st = self._state
        x = melspec.detach()
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        xl = x[:, :, T - 1]
        pma_t = None
        if pma is not None:
            _lib.require_cuda(pma, "melSyn.forward pma")
            pma_t = pma.detach().to(torch.int64).contiguous()
            if pma_t.shape != (B,):
                raise ValueError(f"pma must be ({B},), got {tuple(pma_t.shape)}")
        _lib.check(lib.ssv_decoder_step(dec, xl.data_ptr(), xl.stride(0), xl.stride(1),
                                        None if pma_t is None else pma_t.data_ptr(), _lib.current_stream_ptr()))
        st["t"] = T
        st["last_in"] = (x, pma_t)
        Y, A, max_att = st["Y"][:, :, :T], st["A"][:, :, :T], st["traj"][T - 1]
        if first:
            return Y, A, max_att, K, V
        return Y, A, max_att

    def _train_forward(self, melspec, textid, spkemb):
        """Train branch of the reference (:263-273): teacher-forced full-sequence forward, unmasked attention.

        Returns (Y_prob (B, F, T), A (B, N, T)) like the reference.  Forward only: the outputs carry no autograd
        graph (the backward pass of the training step is the next row of the scope table)."""
        _lib.require_cuda(melspec, "melSyn.forward melspec")
        if textid is None:
            raise ValueError("melSyn.forward in train() mode needs textid")
        _lib.require_cuda(textid, "melSyn.forward textid")
        if melspec.dim() != 3 or melspec.shape[1] != self.freq_bins:
            raise ValueError(f"melspec must be (B, {self.freq_bins}, T), got {tuple(melspec.shape)}")
        ids = textid.detach().to(torch.int64).contiguous()
        if ids.dim() != 3 or ids.shape[1] != 1 or ids.shape[0] != melspec.shape[0]:
            raise ValueError(f"textid must be ({melspec.shape[0]}, 1, N), got {tuple(ids.shape)}")
        B, _, T = melspec.shape
        N = ids.shape[-1]
        if B == 0 or T == 0 or N == 0:
            raise ValueError("empty input")
        x = melspec.detach().to(torch.float32).contiguous()
        spk = spkemb.detach().to(torch.float32).contiguous()
        if spk.shape != (B, self.spkemb_dim, 1):
            raise ValueError(f"spkemb must be ({B}, {self.spkemb_dim}, 1), got {tuple(spk.shape)}")
        Y = torch.empty((B, self.freq_bins, T), device=x.device, dtype=torch.float32)
        A = torch.empty((B, N, T), device=x.device, dtype=torch.float32)
        h = self._native()
        _lib.check(_lib.load().ssv_text2mel_train_fwd(h, x.data_ptr(), ids.data_ptr(), spk.data_ptr(), B, N, T,
                                                      Y.data_ptr(), A.data_ptr(), _prec(self.precision),
                                                      _lib.current_stream_ptr()))
        _lib.check(_lib.load().ssv_text2mel_check(h, _lib.current_stream_ptr()))      # ids verified on the device
        return Y, A

    def _train_forward_autograd(self, melspec, textid, spkemb):
        """Train branch (:263-273) WITH an autograd graph, for `loss.backward()` of the training step
        (train/adversarial_wasserstein_gp.py:277-300).  The 38 highway convs -- 98 % of the TextEnc, 96 % of the
        AudioEnc and 87 % of the AudioDec FLOPs -- run forward and backward in the library's kernels
        (`_HighwayConvFn`), and so do the eleven 1x1 conv + LayerNorm layers (`_ConvLnFn`), the embedding gather, the
        two speaker projections and the unmasked attention (csrc/train_small.cu); what is left to torch is the final
        sigmoid and the graph bookkeeping."""
        import torch.nn.functional as F
        _lib.require_cuda(melspec, "melSyn.forward melspec")
        if textid is None:
            raise ValueError("melSyn.forward in train() mode needs textid")
        te, ae, ad = self.text_encoder, self.audio_encoder, self.audio_decoder

        def cl(x, conv, lnm, relu_in=False, sb=None):       # 1x1 conv (+ speaker term) + LayerNorm, fwd / bwd in the library
            return _ConvLnFn.apply(x, conv.weight, conv.bias, sb, lnm.weight, lnm.bias, relu_in, _prec(self.precision))

        prec = _prec(self.precision)        # "fp32": the highway convs' forward and dgrad on the tensor cores (3xTF32)

        def hc(x, bag, k, dil, causal):                     # x: (B, T, d), see stack()
            return _HighwayConvFn.apply(x, bag.conv.weight, bag.conv.bias, bag.ln1.weight, bag.ln1.bias, bag.ln2.weight,
                                        bag.ln2.bias, k, dil, causal, prec, True)

        def hci(x, bag, causal):
            for i, dil in enumerate((1, 3, 9, 27), start=1):
                x = hc(x, getattr(bag, f"hc{i}"), 3, dil, causal)
            return x

        def stack(x, fn):
            """A run of highway layers on channels-last activations: one transpose in, one out, instead of two per layer
            and direction inside the library calls."""
            return fn(x.transpose(1, 2).contiguous()).transpose(1, 2).contiguous()

        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            ids = textid.long()[:, 0, :]
            x = _TextEmbeddingFn.apply(ids, te.textemb_layer.W.weight, te.textemb_layer.W.bias)
            x = cl(x, te.conv1, te.ln1)
            x = cl(x, te.conv2, te.ln2, relu_in=True)
            def te_stack(x):
                x = hci(hci(x, te.hci1, False), te.hci2, False)
                for bag, k in ((te.hc1, 3), (te.hc2, 3), (te.hc3, 1), (te.hc4, 1)):
                    x = hc(x, bag, k, 1, False)
                return x
            kv = stack(x, te_stack)                         # [K ; V] (B, 2 hidden, N)

            mel = melspec.to(torch.float32)
            spk = spkemb.to(torch.float32)[:, :, 0]
            s1 = _LinearSmallFn.apply(spk, ae.fc1.weight, ae.fc1.bias)
            s2 = _LinearSmallFn.apply(spk, ae.fc2.weight, ae.fc2.bias)
            q = cl(mel, ae.conv1, ae.ln1, sb=s1)
            q = cl(q, ae.conv2, ae.ln2, relu_in=True)
            q = cl(q, ae.conv3, ae.ln3, relu_in=True, sb=s2)
            q = stack(q, lambda z: hc(hc(hci(hci(z, ae.hci1, True), ae.hci2, True), ae.hc1, 3, 3, True), ae.hc2, 3, 3, True))

            A, r = _AttentionTrainFn.apply(kv, q)           # r = [V A ; q]

            y = cl(r, ad.conv1, ad.ln1)
            y = stack(y, lambda z: hc(hc(hci(z, ad.hci, True), ad.hc1, 3, 1, True), ad.hc2, 3, 1, True))
            y = cl(y, ad.conv2, ad.ln2)
            y = cl(y, ad.conv3, ad.ln3, relu_in=True)
            y = cl(y, ad.conv4, ad.ln4, relu_in=True)
            y = cl(y, ad.conv5, ad.ln5, relu_in=True)
            return torch.sigmoid(y), A
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32

    def check(self):
        """Synchronise and raise if the decode kernel aborted."""
        if self._dec is not None:
            _lib.check(_lib.load().ssv_decoder_check(C.c_void_p(self._dec), _lib.current_stream_ptr()))


class SSRN(_Native):
    """reference models/TTSModel.py:319-362."""

    _create_name = "ssv_ssrn_create"
    _destroy_name = "ssv_ssrn_destroy"

    def __init__(self, freq_bins, output_bins, ssrn_dim):
        super().__init__()
        self.freq_bins, self.output_bins, self.ssrn_dim = freq_bins, output_bins, ssrn_dim
        d, d2, o = ssrn_dim, 2 * ssrn_dim, output_bins
        pw = lambda i, o_: nn.Conv1d(i, o_, 1)
        ups = lambda: _Bag(deconv=nn.ConvTranspose1d(d, d, 2, stride=2), hc1=_hc(d, 3), hc2=_hc(d, 3))
        for name, mod in (
                ("conv1", pw(freq_bins, d)), ("ln1", nn.LayerNorm(d)), ("hc1", _hc(d, 3)), ("hc2", _hc(d, 3)),
                ("ups1", ups()), ("ups2", ups()),
                ("conv2", pw(d, d2)), ("ln2", nn.LayerNorm(d2)), ("hc3", _hc(d2, 3)), ("hc4", _hc(d2, 3)),
                ("conv3", pw(d2, o)), ("ln3", nn.LayerNorm(o)), ("conv4", pw(o, o)), ("ln4", nn.LayerNorm(o)),
                ("conv5", pw(o, o)), ("ln5", nn.LayerNorm(o)), ("conv6", pw(o, o)), ("ln6", nn.LayerNorm(o))):
            self.add_module(name, mod)
        self.precision = "fp32"

    def _dims(self):
        return (self.freq_bins, self.output_bins, self.ssrn_dim)

    def _forward_autograd(self, inputs):
        """SSRN.forward (:342-362) with an autograd graph, for the `train_ssrn` step
        (train/adversarial_wasserstein_gp.py:324-360): the eight highway convs (94 % of the FLOPs) forward and
        backward in the library's kernels, the 1x1 convs, the two transposed convs and the LayerNorms as FP32 torch
        matmuls / ops."""
        import torch.nn.functional as F

        def ln(x, m):
            return F.layer_norm(x.transpose(1, 2), (m.weight.numel(),), m.weight, m.bias, 1e-5).transpose(1, 2)

        def pw(x, m):
            return torch.matmul(m.weight[:, :, 0], x) + m.bias[None, :, None]

        prec = _prec(self.precision)

        def hc(x, bag, dil):
            return _HighwayConvFn.apply(x, bag.conv.weight, bag.conv.bias, bag.ln1.weight, bag.ln1.bias, bag.ln2.weight,
                                        bag.ln2.bias, 3, dil, False, prec)

        def ups(x, bag):
            w = bag.deconv.weight                                   # (in, out, 2), stride 2: out[2t + j] = W[:, :, j]^T x[t]
            y = torch.stack((torch.matmul(w[:, :, 0].t(), x), torch.matmul(w[:, :, 1].t(), x)), dim=-1)
            y = y.reshape(x.shape[0], w.shape[1], 2 * x.shape[2]) + bag.deconv.bias[None, :, None]
            return hc(hc(y, bag.hc1, 1), bag.hc2, 3)

        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            x = ln(pw(inputs.to(torch.float32), self.conv1), self.ln1)
            x = hc(hc(x, self.hc1, 1), self.hc2, 3)
            x = ups(ups(x, self.ups1), self.ups2)
            x = ln(pw(x, self.conv2), self.ln2)
            x = hc(hc(x, self.hc3, 1), self.hc4, 1)
            x = ln(pw(x, self.conv3), self.ln3)
            x = ln(pw(x, self.conv4), self.ln4)
            x = ln(pw(F.relu(x), self.conv5), self.ln5)
            x = ln(pw(F.relu(x), self.conv6), self.ln6)
            return torch.sigmoid(x)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32

    def forward(self, inputs):
        _lib.require_cuda(inputs, "SSRN.forward")
        if inputs.dim() != 3 or inputs.shape[1] != self.freq_bins:
            raise ValueError(f"SSRN expects (B, {self.freq_bins}, T), got {tuple(inputs.shape)}")
        if (self.training and self.precision in ("fp32", "fp32-ffma") and torch.is_grad_enabled()
                and any(p.requires_grad for p in self.parameters()) and inputs.numel() > 0):
            return self._forward_autograd(inputs)
        x = inputs.detach()
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        B, _, T = x.shape
        out = torch.empty((B, self.output_bins, 4 * T), device=x.device, dtype=torch.float32)
        if B == 0 or T == 0:
            return out
        _lib.check(_lib.load().ssv_ssrn_fwd(self._native(), x.data_ptr(), x.stride(0), x.stride(1), x.stride(2),
                                            B, T, out.data_ptr(), _prec(self.precision), _lib.current_stream_ptr()))
        return out
