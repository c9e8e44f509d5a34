"""Synthesis driver: whole-utterance batches and utterance sharding across GPUs.

Mirrors the hot part of generate_test_utterances.py:98-124 (per speaker: AR loop over a padded
batch of sentences, SSRN, device->host copy).  Work units (speaker, sentence) are independent, so
multi-GPU is one process per GPU taking a contiguous slice of the units -- no collective on the
data path.  Griffin-Lim / wav writing (generate_test_utterances.py:126-139) is the
next row of the scope table and is not done here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .models.TTSModel import SSRN, _prec, melSyn


def shard_range(n_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of n_units for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_units, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Unit:
    speaker: int     # index into the speaker list
    sentence: int    # index into the sentence list


def corpus_units(n_speakers: int, n_sentences: int) -> List[Unit]:
    """Speaker-major enumeration, the order generate_test_utterances.py:100-139 writes files in."""
    return [Unit(s, k) for s in range(n_speakers) for k in range(n_sentences)]


def plan_batches(units: Sequence[Unit], batch: int) -> List[List[Unit]]:
    return [list(units[i:i + batch]) for i in range(0, len(units), batch)]


class Synthesizer:
    """Text2Mel + SSRN on one GPU through the C ABI's host-buffer entry point."""

    def __init__(self, text2mel: melSyn, ssrn: SSRN, ssrn_precision: str = "fp32", lin_dtype: str = "fp32"):
        """lin_dtype: element type of the returned linear spectrogram, "fp32" (the reference's) or "bf16" (half the
        device->host bytes; `out["lin"]` is then a torch.bfloat16 CPU tensor instead of a numpy array)."""
        if lin_dtype not in ("fp32", "bf16"):
            raise ValueError("lin_dtype must be 'fp32' or 'bf16'")
        self.m1, self.m2 = text2mel, ssrn
        self.ssrn_precision = ssrn_precision
        self.lin_dtype = lin_dtype
        self._pinned: Dict[str, torch.Tensor] = {}
        self._next_buf = 0

    def _pin(self, key: str, shape, dtype) -> torch.Tensor:
        t = self._pinned.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self._pinned[key] = t
        return t

    def synthesize_host(self, textid: np.ndarray, spkemb: np.ndarray, n_frames: int, want_mel: bool = False,
                        want_att: bool = False) -> Dict[str, np.ndarray]:
        """textid (B, N) int64 and spkemb (B, E) fp32 on the HOST -> host arrays.

        Everything between the host buffers (H2D, TextEnc, n_frames of decode, SSRN, D2H, sync)
        happens inside one C call, ssv_synthesize_host."""
        return self.collect(self.submit(textid, spkemb, n_frames, want_mel, want_att, _sync=True))

    def submit(self, textid: np.ndarray, spkemb: np.ndarray, n_frames: int, want_mel: bool = False,
               want_att: bool = False, _sync: bool = False):
        """Pipelined form (ssv_synthesize_host_submit): enqueue one batch and return a ticket for `collect`.

        Two batches may be in flight: the device->host copy of batch i then overlaps TextEnc / decode of batch
        i + 1.  Each in-flight batch has its own pinned host buffers, which `collect` returns (valid until the
        second-next submit)."""
        ids = np.ascontiguousarray(textid, dtype=np.int64)
        spk = np.ascontiguousarray(spkemb, dtype=np.float32)
        if ids.ndim != 2 or spk.ndim != 2 or ids.shape[0] != spk.shape[0]:
            raise ValueError("textid must be (B, N) and spkemb (B, E)")
        if ids.size and (ids.min() < 0 or ids.max() >= self.m1.vocab_len):
            raise ValueError(f"text ids must lie in [0, {self.m1.vocab_len})")
        B, N = ids.shape
        T = int(n_frames)
        F, O = self.m1.freq_bins, self.m2.output_bins
        k = self._next_buf                                  # host buffer set of this batch (two sets alternate)
        self._next_buf ^= 1
        h_ids = self._pin(f"ids{k}", (B, N), torch.int64)
        h_spk = self._pin(f"spk{k}", (B, spk.shape[1]), torch.float32)
        h_ids.numpy()[...] = ids
        h_spk.numpy()[...] = spk
        lin16 = self.lin_dtype == "bf16"
        h_lin = self._pin(f"lin{k}", (B, O, 4 * T), torch.bfloat16 if lin16 else torch.float32)
        h_traj = self._pin(f"traj{k}", (T, B), torch.int64)
        h_mel = self._pin(f"mel{k}", (B, F, T), torch.float32) if want_mel else None
        h_att = self._pin(f"att{k}", (B, N, T), torch.float32) if want_att else None
        dec = self.m1._decoder(B, N, T)
        _lib.check(_lib.load().ssv_decoder_set_lin_output(dec, int(lin16)))      # no-op unless it changes
        ptr = lambda t: None if t is None else t.data_ptr()
        ticket = C.c_int(-1)
        args = (self.m1._native(), dec, self.m2._native(), h_ids.data_ptr(), h_spk.data_ptr(), B, N, T,
                h_lin.data_ptr(), ptr(h_mel), ptr(h_att), h_traj.data_ptr(),
                _prec(self.m1.precision), _prec(self.ssrn_precision), _lib.current_stream_ptr())
        if _sync:                   # the one-call form: returns with the host buffers filled
            _lib.check(_lib.load().ssv_synthesize_host(*args))
        else:
            _lib.check(_lib.load().ssv_synthesize_host_submit(*args, C.byref(ticket)))
        self.m1._state = None       # the decoder's buffers now belong to the C side
        out = {"lin": h_lin if lin16 else h_lin.numpy(), "traj": h_traj.numpy()}
        if want_mel:
            out["mel"] = h_mel.numpy()
        if want_att:
            out["att"] = h_att.numpy()
        self.h2d_bytes = ids.nbytes + spk.nbytes
        self.d2h_bytes = h_lin.numel() * h_lin.element_size() + h_traj.numel() * 8 + (h_mel.numel() * 4 if want_mel else 0) + (
            h_att.numel() * 4 if want_att else 0)
        return (dec, ticket.value, out)

    def collect(self, pending) -> Dict[str, np.ndarray]:
        """Wait for a submitted batch (ssv_synthesize_host_wait) and return its host arrays."""
        dec, ticket, out = pending
        if ticket >= 0:
            _lib.check(_lib.load().ssv_synthesize_host_wait(dec, ticket))
        return out

    def synthesize_units(self, units: Sequence[Unit], sentence_ids: Sequence[np.ndarray], speaker_emb: np.ndarray,
                         n_frames: int, batch: int, pad_to: Optional[int] = None):
        """Yield (units_of_batch, host outputs) for a rank's slice of the corpus.

        All batches are padded to the same N (default: the longest sentence of the whole list), as
        the reference pads every sentence of its batch to the batch maximum."""
        from .text import pad_batch
        n = pad_to or max(int(np.asarray(r).size) for r in sentence_ids)
        prev = None                      # one batch ahead: batch i + 1 computes while batch i copies back
        for group in plan_batches(units, batch):
            ids = pad_batch([sentence_ids[u.sentence] for u in group], n)
            spk = np.stack([speaker_emb[u.speaker] for u in group]).astype(np.float32)
            if prev is not None and prev[2] != ids.shape:       # shape change (last, smaller batch): drain first
                yield prev[0], self.collect(prev[1])
                prev = None
            cur = (group, self.submit(ids, spk, n_frames), ids.shape)
            if prev is not None:
                yield prev[0], self.collect(prev[1])
            prev = cur
        if prev is not None:
            yield prev[0], self.collect(prev[1])
