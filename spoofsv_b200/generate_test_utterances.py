"""CLI-compatible driver for the corpus synthesis of the reference's generate_test_utterances.py.

    python -m spoofsv_b200.generate_test_utterances -C config.json -T <time tag> [--eval_utt_num 20]
    python -m torch.distributed.run --nproc-per-node 8 -m spoofsv_b200.generate_test_utterances -C ... -T ...

Same flags and the same config.json keys as the reference (generate_test_utterances.py:44-93): the first
`eval_utt_num` lines of TTS_TEXTS are padded to their common maximum and synthesised for every speaker of
SPK_EMB_DIR for MAX_FRAME_NUM + 1 frames (:105-116), then SSRN (:120).  What is different:

  * TextEnc runs ONCE for the sentence batch; its K/V are gathered per (speaker, sentence) unit -- they do not
    depend on the speaker (the reference recomputes them per speaker, :108);
  * units are batched `--batch` at a time irrespective of speaker boundaries and, under torchrun, sharded
    contiguously over the ranks (one process per GPU, no collective on the data path);
  * the AR loop is one launch of the incremental decode kernel per batch.

With --save_spectrogram DIR it writes the linear spectrograms as `s<spk>/s<spk>_<nnn>.npy`; with --save_wav DIR it
also runs the waveform stage of the reference (:126-139: Griffin-Lim, de-emphasis, trim, wav files) batched on
the GPU (spoofsv_b200/vocoder.py) and writes `s<spk>/s<spk>_<nnn>.wav`; --write_layouts then lays them out for the
Kaldi / GE2E / anti-spoofing consumers as :141-259 does (spoofsv_b200/protocols.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from . import text as T
from .synth import Unit, plan_batches, shard_range


def build_parser() -> argparse.ArgumentParser:
    ps = argparse.ArgumentParser(description="Adversarial Conditional Text-to-speech")
    ps.add_argument("-C", "--configuration", type=str, default=None)
    ps.add_argument("--train_spk_num", type=int, default=88)      # used by --write_layouts (the Kaldi / GE2E file
    ps.add_argument("--enroll_utt_num", type=int, default=3)      # shuffling of the reference) only
    ps.add_argument("--eval_utt_num", type=int, default=20)
    ps.add_argument("-T", "--current_time", type=str, required=True)
    # extensions
    ps.add_argument("--batch", type=int, default=64, help="utterances per decode launch")
    ps.add_argument("--frames", type=int, default=None, help="frames per utterance (default MAX_FRAME_NUM + 1)")
    ps.add_argument("--speakers", type=int, default=None, help="only the first N speakers (dry runs)")
    ps.add_argument("--ssrn_precision", default="fp32", choices=["fp32", "bf16"],
                    help="fp32 (default, the reference's numerics: 1e-4 max-abs) or bf16 (tcgen05 tensor cores, 2e-2 "
                         "relative L2 on the linear spectrogram; opt-in)")
    ps.add_argument("--save_spectrogram", type=str, default=None, help="directory for s<spk>/s<spk>_<nnn>.npy")
    ps.add_argument("--save_wav", type=str, default=None,
                    help="directory for s<spk>/s<spk>_<nnn>.wav: batched GPU Griffin-Lim (64 iterations), de-emphasis, "
                         "trim, 9 s cap, peak 0.75, as generate_test_utterances.py:130-139 (default there: "
                         "SRC_ROOT_DIR/test/<current_time>/spoof_data/)")
    ps.add_argument("--gl_iters", type=int, default=64)
    ps.add_argument("--write_layouts", action="store_true",
                    help="after synthesis (rank 0, needs --save_wav and the VCTK corpus under DATA_ROOT_DIR): the Kaldi "
                         "i-vector tree, GE2E links and, when ANTISPOOF_DIR exists, the anti-spoofing set of "
                         "generate_test_utterances.py:141-259 under SRC_ROOT_DIR/test/<current_time>/")
    ps.add_argument("--random_init", type=int, default=None, metavar="SEED",
                    help="random-init weights instead of the INFERENCE_* checkpoints (no checkpoints are vendored)")
    return ps


def load_config(path: str) -> dict:
    with open(path, "r") as f:
        return json.load(f)


def read_texts(cfg: dict, n: int) -> np.ndarray:
    """First n lines of TTS_TEXTS -> ids padded to their common maximum, as generate_test_utterances.py:56-73."""
    with open(cfg["TTS_TEXTS"], "r") as f:
        lines = [s.strip() for s in f.readlines()]
    if len(lines) < n:
        raise ValueError(f"TTS_TEXTS holds {len(lines)} lines, eval_utt_num={n}")
    return T.pad_batch(T.encode_lines(lines[:n], cfg["VOCABULARY"]))


def list_speakers(cfg: dict) -> List[str]:
    """Speaker directory names of DATA_ROOT_DIR/wav22 (reference :99) or, without the corpus, the stems of
    SPK_EMB_DIR/*.npy."""
    wav = os.path.join(cfg.get("DATA_ROOT_DIR", ""), "wav22")
    if os.path.isdir(wav):
        return sorted(os.listdir(wav))
    return sorted(p.stem for p in Path(cfg["SPK_EMB_DIR"]).glob("*.npy"))


def plan(n_speakers: int, n_sentences: int, world: int, rank: int, batch: int) -> List[List[Unit]]:
    """This rank's batches of (speaker, sentence) units, speaker-major like the reference's output order."""
    lo, hi = shard_range(n_speakers * n_sentences, world, rank)
    units = [Unit(i // n_sentences, i % n_sentences) for i in range(lo, hi)]
    return plan_batches(units, batch)


def output_name(spk: str, sentence: int) -> str:
    """s<spk>/s<spk>_<nnn> -- the reference's wav naming (:139) without the extension."""
    return f"s{spk[1:]}/s{spk[1:]}_{str(sentence + 1).zfill(3)}"


class _HostPipe:
    """Device -> host copies of the spectrogram batches on a side stream into two alternating pinned buffers."""

    def __init__(self):
        import torch
        self.torch = torch
        self.stream = torch.cuda.Stream()
        self.bufs = [None, None]
        self.k = 0

    def submit(self, t):
        torch = self.torch
        k = self.k
        self.k ^= 1
        if self.bufs[k] is None or self.bufs[k].shape != t.shape:
            self.bufs[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record()                                   # the producing kernels on the current stream
        done = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.bufs[k].copy_(t, non_blocking=True)
            done.record()
        t.record_stream(self.stream)                     # the allocator must not hand the device buffer out before the copy ran
        return k, done

    def collect(self, ticket):
        k, done = ticket
        done.synchronize()
        return self.bufs[k].numpy()                      # valid until the second-next submit


def run(args, on_batch=None) -> Dict[str, float]:
    import torch
    from .models.TTSModel import SSRN, melSyn
    from . import _lib

    cfg = load_config(args.configuration)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("generate_test_utterances: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)

    text_id = read_texts(cfg, args.eval_utt_num)                       # (U, N)
    speakers = list_speakers(cfg)
    if args.speakers is not None:
        speakers = speakers[: args.speakers]
    frames = args.frames or cfg["MAX_FRAME_NUM"] + 1

    if args.random_init is not None:
        torch.manual_seed(args.random_init)
    m1 = melSyn(vocab_len=len(cfg["VOCABULARY"]) - 1, condition=True, spkemb_dim=cfg["SPK_EMB_DIM"],
                textemb_dim=cfg["TEXT_EMB_DIM"], freq_bins=cfg["COARSE_MELSPEC"]["FREQ_BINS"],
                hidden_dim=cfg["HIDDEN_DIM"])
    m2 = SSRN(freq_bins=cfg["COARSE_MELSPEC"]["FREQ_BINS"], output_bins=1 + cfg["STFT"]["FFT_LENGTH"] // 2,
              ssrn_dim=cfg["SSRN_DIM"])
    if args.random_init is None:
        m1.load_state_dict(torch.load(cfg["INFERENCE_TEXT2MEL_MODEL"], map_location="cpu")["model_state_dict"])
        m2.load_state_dict(torch.load(cfg["INFERENCE_SSRN_MODEL"], map_location="cpu")["model_state_dict"])
    m1, m2 = m1.cuda().eval(), m2.cuda().eval()
    m2.precision = args.ssrn_precision

    emb = np.stack([np.load(os.path.join(cfg["SPK_EMB_DIR"], s + ".npy")).astype(np.float32) for s in speakers])
    emb_d = torch.from_numpy(emb).cuda()
    K, V = m1.encode_text(torch.from_numpy(text_id)[:, None, :].cuda())    # once: K/V do not depend on the speaker
    lib = _lib.load()
    batches = plan(len(speakers), text_id.shape[0], world, rank, args.batch)
    want_host = on_batch is not None or bool(args.save_spectrogram)
    pipe = _HostPipe() if want_host else None
    n_utt, t0 = 0, time.perf_counter()

    def finish(job):
        """Second half of a batch, one batch behind the launches: host copy done -> callbacks and files."""
        group, ticket, state = job
        lin = pipe.collect(ticket)
        if on_batch is not None:
            on_batch(group, lin, state)
        if args.save_spectrogram:
            for u, spec in zip(group, lin):
                path = Path(args.save_spectrogram) / (output_name(speakers[u.speaker], u.sentence) + ".npy")
                path.parent.mkdir(parents=True, exist_ok=True)
                np.save(path, spec)

    pending = None
    with torch.no_grad():
        for group in batches:
            sent = torch.tensor([u.sentence for u in group], device="cuda")
            spk = emb_d[torch.tensor([u.speaker for u in group], device="cuda")][:, :, None]
            dec = m1._begin(K[sent].contiguous(), V[sent].contiguous(), spk, frames)
            _lib.check(lib.ssv_decoder_run(dec, frames, _lib.current_stream_ptr()))
            lin_d = m2(m1._state["Y"])                                      # (B, 513, 4T), the reference's pred_lin
            # the spectrogram batch (114 MB at 64 utterances) crosses PCIe only when somebody wants it on the host, and
            # then on a copy stream into one of two pinned buffers, under the next batch's decode (the same overlap
            # as ssv_synthesize_host_submit / _wait, with the text encoder hoisted out of the loop)
            job = None
            if want_host:
                st = m1._state
                job = (group, pipe.submit(lin_d), {"Y": st["Y"].clone(), "A": st["A"].clone(), "traj": st["traj"].clone()})
            if args.save_wav:
                from . import vocoder
                waves = vocoder.postprocess(lin_d, cfg, n_iter=args.gl_iters)
                for u, w in zip(group, waves):
                    path = Path(args.save_wav) / (output_name(speakers[u.speaker], u.sentence) + ".wav")
                    path.parent.mkdir(parents=True, exist_ok=True)
                    vocoder.write_wav(path, w, cfg["SAMPLING_RATE"])
            if pending is not None:
                finish(pending)
            pending = job
            n_utt += len(group)
        if pending is not None:
            finish(pending)
        m1.check()
    sec = time.perf_counter() - t0
    stats = {"rank": rank, "world": world, "utterances": n_utt, "frames": frames, "seconds": sec,
             "frames_per_s": n_utt * frames / sec if sec > 0 else 0.0}
    if getattr(args, "write_layouts", False):
        if world > 1:
            _wait_for_all_ranks(args, rank, world)   # every rank's wavs are on disk before rank 0 reads them
        if rank == 0:
            stats["layouts"] = write_layouts(args, cfg)
    print(json.dumps(stats), flush=True)
    return stats


def _wait_for_all_ranks(args, rank: int, world: int, timeout_s: float = 24 * 3600.0, poll_s: float = 0.5) -> None:
    """Synthesis uses no collective (and no process group), so the ranks meet through the file system: every rank
    drops `<save_wav>/.done.<current_time>.<rank>` when its shard is on disk and rank 0 waits for all of them."""
    if not args.save_wav:
        raise ValueError("--write_layouts needs --save_wav (the synthesised wavs are its input)")
    root = Path(args.save_wav)
    root.mkdir(parents=True, exist_ok=True)
    mark = lambda r: root / ".done.{}.{}".format(args.current_time, r)
    mark(rank).write_text("ok\n")
    if rank != 0:
        return
    t0 = time.perf_counter()
    while not all(mark(r).exists() for r in range(world)):
        if time.perf_counter() - t0 > timeout_s:
            missing = [r for r in range(world) if not mark(r).exists()]
            raise RuntimeError("write_layouts: ranks {} did not finish their shard".format(missing))
        time.sleep(poll_s)
    for r in range(world):
        mark(r).unlink()


def write_layouts(args, cfg: dict) -> dict:
    """generate_test_utterances.py:141-259 for the wavs under --save_wav."""
    from . import protocols as P
    if not args.save_wav:
        raise ValueError("--write_layouts needs --save_wav (the synthesised wavs are its input)")
    with open(cfg["TTS_TEXTS"], "r") as f:
        sentences = [s.strip() for s in f.readlines()]
    out_root = os.path.join(cfg["SRC_ROOT_DIR"], "test", args.current_time)
    done = {"ivector": P.write_ivector_layout(cfg["DATA_ROOT_DIR"], args.save_wav, out_root, sentences, args.train_spk_num,
                                              args.enroll_utt_num, args.eval_utt_num),
            "ge2e_links": P.link_ge2e(out_root)}
    anti = cfg.get("ANTISPOOF_DIR")
    if anti and os.path.isdir(os.path.join(anti, "ASVspoof2019_LA_cm_protocols")):
        done["antispoof"] = P.write_antispoof_set(anti, args.save_wav, args.current_time)
    return done


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    if args.configuration is None:
        print("generate_test_utterances: -C/--configuration is required", file=sys.stderr)
        return 2
    run(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
