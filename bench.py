"""Benchmark of the Text2Mel + SSRN synthesis hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic utterances shaped like
generate_test_utterances.py's workload (BASELINE config 4 = configs 2+3 back to back): B utterances
(Harvard sentences x VCTK speaker embeddings, all padded to N = 58), TextEnc once, T = 217 frames
of incremental decode, SSRN on the result.  Frames are REDUCED MEL FRAMES (decode steps; one mel
frame = 4 linear-spectrogram frames).

  value     frames/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e       the same through the C ABI's host-buffer call (ssv_synthesize_host): H2D of ids +
            embeddings and D2H of the linear spectrogram inside the timed region
  roofline  the dominant kernel (decode_ws_kernel, the weight-stationary pipelined decode), algorithmic bytes /
            CUDA-event time against the measured HBM copy bandwidth (MEASURED_PEAKS.json); roofline_ssrn: the
            tcgen05 conv stack against the measured sustained bf16 peak
  cpu_baseline / --impl reference
            the reference algorithm (re-encoding O(T^2) AR loop + SSRN) as restated in oracle/ on the
            host cores, on a bounded sample of the same workload.  /root/reference is a Python repo
            that does not exist on the GPU box, so the port is what can be timed there.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Text2Mel+SSRN synthesized frames/s"
UNIT = "reduced mel frames/s"
N_TEXT = 58
T_FRAMES = 217
BATCH = 256            # utterances per step and GPU (decode throughput: 1.29 M frames/s at 64, 1.43 M at 128, 1.55 M at 256)
# SURVEY.md 8(d) cfg 3: algorithmic bytes per decode step = AudioEnc+AudioDec weights (fc1/fc2 hoisted)
DECODE_WEIGHT_BYTES = 6_821_360 * 4
DECODE_STATE_BYTES_PER_UTT = 16 * 3 * 256 * 4 + 6 * 1024 + 640       # taps r/w, K/V window, x/y  (~56 KB)
SSRN_FLOP_PER_FRAME = 46_469_144
TEXTENC_FLOP_PER_CHAR = 34_218_496
DECODE_FLOP_PER_FRAME = 13_630_000
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12         # 74.45


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 30 for the native arm, 8 for the CPU reference arm: 3.4 s per step there)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--ssrn-precision", default="bf16", choices=["fp32", "bf16"],
                    help="SSRN arm: bf16 = tcgen05 tensor cores (2e-2 rel-L2 bar), fp32 = 3xTF32 on the tensor cores (1e-4 max-abs bar); "
                         "Text2Mel always runs the fp32 arm (identical alignments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=4, help="utterances in the CPU baseline sample")
    ap.add_argument("--dump-lin", default=None, help="(reference arm) write the sample's linear spectrogram to this .npy")
    ap.add_argument("--no-eager", action="store_true", help="skip extra.torch_eager_gpu")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 8 if args.impl == "reference" else 30
    return args


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons sampled every 100 ms while the timed region runs.

    In-process NVML (nvidia_ml_py): a query is a few microseconds of driver work.  A looping `nvidia-smi` child
    (the fallback when NVML cannot be imported) stalled kernel launches for ~160 ms whenever a sample fell into a
    step, which showed up as one 180 ms step among 17 ms ones."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid, self.rows, self.proc = index, uuid, [], None
        self.nvml, self.handle, self.stop = None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            self.handle = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            # first use of a query initialises driver state (tens of ms, seen as one slow step): do it here
            for _ in range(2):
                pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = [(n.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                (n.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]
        while not self.stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append([sm, self.max_sm, [name for bit, name in bits if mask & bit]])
            except Exception:
                pass
            self.stop.wait(0.1)

    def _read_smi(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append([float(c[0]), float(c[1]),
                                  [n for n, v in zip(self.NAMES, c[3:7]) if v.lower().startswith("active")]])
            except (ValueError, IndexError):
                continue

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [r[0] for r in self.rows]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(r[1] for r in self.rows), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def workload(batch: int, rank: int):
    """Rank r takes units [r*B, (r+1)*B) of the speaker-major corpus (108 speakers x 720 sentences)."""
    import numpy as np
    from oracle import weights as W
    from spoofsv_b200.synth import corpus_units
    from spoofsv_b200.text import encode_lines, pad_batch
    names, emb, lines = W.load_fixtures()
    sent = encode_lines(lines)
    units = corpus_units(len(names), len(lines))
    # stride through the corpus so a batch mixes speakers and sentences
    stride = 1201
    pick = [units[((rank * batch + i) * stride) % len(units)] for i in range(batch)]
    ids = pad_batch([sent[u.sentence] for u in pick], N_TEXT)
    spk = np.stack([emb[u.speaker] for u in pick]).astype(np.float32)
    return ids, spk


def cpu_reference_step(sd1, sd2, ids, spk, frames):
    """One step of the reference algorithm on the CPU (oracle port): re-encoding AR loop + SSRN."""
    import torch
    from oracle import ttsmodel_oracle as O
    with torch.no_grad():
        t0 = time.perf_counter()
        Y, A, traj, lin = O.synthesize(sd1, sd2, torch.from_numpy(ids)[:, None, :], torch.from_numpy(spk)[:, :, None], frames)
        return time.perf_counter() - t0, lin


def reference_module_step(R, r1, r2, ids, spk, frames):
    """One step through the UNMODIFIED reference module (oracle/_ref/models/TTSModel.py) on its own `device`:
    the AR loop of generate_test_utterances.py:105-116 + SSRN (:120)."""
    import torch
    from oracle import build_ref
    with torch.no_grad():
        tid = torch.from_numpy(ids)[:, None, :].to(R.device)
        e = torch.from_numpy(spk)[:, :, None].to(R.device)
        if R.device.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        Y, A, pma = build_ref.ar_loop(R, r1, tid, e, frames)
        lin = r2(Y)
        if R.device.type == "cuda":
            torch.cuda.synchronize()
        return time.perf_counter() - t0, lin


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, rank 0 only.

    With oracle/_ref staged (oracle/build_ref.py) this is the unmodified models/TTSModel.py (`kind: "reference"`),
    otherwise the oracle port (`kind: "port"`).  Every step is a bounded sample of the native arm's workload
    (`--cpu-sample` of its `batch_per_gpu` utterances, all frames); `config` is the native arm's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""            # the reference binds `device` at import: this arm is the CPU path
    import warnings
    warnings.filterwarnings("ignore")
    import numpy as np
    import torch
    from oracle import build_ref
    from oracle import weights as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd1, sd2 = W.state_dicts(0)
    nb = max(1, min(args.cpu_sample, args.batch))
    ids, spk = workload(args.batch, 0)
    ids, spk = ids[:nb], spk[:nb]
    if build_ref.available():
        R = build_ref.load()
        r1, r2 = build_ref.models(R, sd1, sd2, W.CFG)
        step = lambda frames: reference_module_step(R, r1, r2, ids, spk, frames)
        kind, what = "reference", "unmodified reference models/TTSModel.py (oracle/_ref) driven by the loop of generate_test_utterances.py:105-116"
    else:
        step = lambda frames: cpu_reference_step(sd1, sd2, ids, spk, frames)
        kind, what = "port", "oracle port of models/TTSModel.py"
    for _ in range(args.warmup):
        step(min(args.frames, 8))       # short warm-up: thread pools, oneDNN primitives
    times, lin = [], None
    for _ in range(args.steps):
        sec, lin = step(args.frames)
        times.append(sec)
    if args.dump_lin:
        np.save(args.dump_lin, lin.cpu().numpy())
    total = sum(times)
    value = nb * args.frames * args.steps / total
    sample = (f"{nb} of the {args.batch} utterances per step, all {args.frames} frames, re-encoding AR loop + SSRN: "
              f"{what}, torch {torch.__version__} CPU fp32, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config_dict(args, args.batch),
        "bounded_sample": {"utterances_per_step": nb, "of_batch_per_gpu": args.batch,
                           "note": "ms_per_step is the time of the sample; value = sample frames / sample time"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def torch_eager_gpu(sd1, sd2, ids, spk, frames):
    """Same-GPU comparison, outside the timed region, clearly NOT the product path: stock eager PyTorch on this B200.
    (a) the unmodified reference module (oracle/_ref) driven by the reference's re-encoding AR loop + SSRN, whole
    batch; (b) the oracle's O(T) incremental loop + SSRN as eager torch ops on the GPU.  PyTorch defaults (cuDNN convs
    may use TF32).  One warm-up of a few frames, then one timed run each (wall clock around synchronize)."""
    import torch
    from oracle import build_ref
    from oracle import ttsmodel_oracle as O
    from oracle import weights as W
    out = {"note": "stock eager PyTorch on the same GPU, cudnn.allow_tf32 default; reported for context, not the product path"}
    B = ids.shape[0]
    try:
        with torch.no_grad():
            if build_ref.available():
                R = build_ref.load()
                if R.device.type == "cuda":
                    r1, r2 = build_ref.models(R, sd1, sd2, W.CFG)
                    reference_module_step(R, r1, r2, ids, spk, 6)
                    sec, _ = reference_module_step(R, r1, r2, ids, spk, frames)
                    out["reference_loop"] = {"ms": 1e3 * sec, "frames_per_s": B * frames / sec, "batch": B,
                                             "what": "unmodified models/TTSModel.py, generate_test_utterances.py:105-116 loop + SSRN"}
                    del r1, r2
            g1 = {k: v.cuda() for k, v in sd1.items()}
            g2 = {k: v.cuda() for k, v in sd2.items()}
            tid = torch.from_numpy(ids)[:, None, :].cuda()
            e = torch.from_numpy(spk)[:, :, None].cuda()
            with torch.device("cuda"):
                O.synthesize(g1, g2, tid, e, 6, incremental=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                O.synthesize(g1, g2, tid, e, frames, incremental=True)
                torch.cuda.synchronize()
                sec = time.perf_counter() - t0
            out["incremental_loop"] = {"ms": 1e3 * sec, "frames_per_s": B * frames / sec, "batch": B,
                                       "what": "oracle incremental (O(T)) loop + SSRN as eager torch ops on the GPU"}
    except Exception as exc:      # context only: never fail the bench line over it
        out["error"] = f"{type(exc).__name__}: {exc}"
    return out


def config_dict(args, batch):
    return {"workload": f"generate_test_utterances batch: {batch} utterances (Harvard sentences x VCTK speaker embeddings, "
                        f"N={N_TEXT} padded ids), TextEnc + {args.frames}-frame incremental Text2Mel decode + SSRN "
                        f"-> ({batch}, 513, {4 * args.frames})",
            "batch_per_gpu": batch, "text_len": N_TEXT, "frames": args.frames, "weights": "random-init seed 0",
            "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"utterance-sharded x{args.gpus}, no collective"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import weights as W
    from spoofsv_b200 import _lib
    from spoofsv_b200.synth import Synthesizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = _lib.load()
    B, T = args.batch, args.frames
    m1, m2 = W.build_models(0)
    sd1 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    sd2 = {k: v.detach().clone() for k, v in m2.state_dict().items()}
    m1, m2 = m1.cuda(), m2.cuda()
    m2.precision = args.ssrn_precision
    ids_np, spk_np = workload(B, rank)
    ids_d = torch.from_numpy(ids_np)[:, None, :].cuda()
    spk_d = torch.from_numpy(spk_np)[:, :, None].cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # end-to-end arm: the spectrogram leaves the device in the SSRN arm's own operand precision (bf16 arm: bf16 host
    # buffer, half the D2H bytes; fp32 arm: fp32); the fp32-output variant of the bf16 arm is reported beside it
    lin_dtype = "bf16" if args.ssrn_precision == "bf16" else "fp32"
    syn = Synthesizer(m1, m2, ssrn_precision=args.ssrn_precision, lin_dtype=lin_dtype)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def device_step(parts=None, ids_x=None, spk_x=None):
        ids_x = ids_d if ids_x is None else ids_x
        spk_x = spk_d if spk_x is None else spk_x
        e = [ev() for _ in range(4)]
        e[0].record()
        K, V = m1.encode_text(ids_x, check=False)      # ids are verified on the device; m1.check() after the loop reports
        dec = m1._begin(K, V, spk_x, T)
        e[1].record()
        _lib.check(lib.ssv_decoder_run(dec, T, _lib.current_stream_ptr()))
        e[2].record()
        lin = m2(m1._state["Y"])
        e[3].record()
        return e, lin

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid)          # NVML is initialised here, outside the timed region
    sampler.__enter__()                              # the sampling thread starts before the warm-up steps, so whatever
                                                     # its first queries initialise is not inside the timed region

    # ---- warm-up (also builds the native handles and the decoder)
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        device_step()
        syn.synthesize_host(ids_np, spk_np, T)
    # the timed loop keeps the previous step's result alive while the next one is produced: reach that allocator state
    # here (the second live 114 MB result was otherwise cudaMalloc'ed inside timed step 1: one 25-200 ms step)
    for _ in range(3):
        flush.fill_(1)
        e, lin = device_step()
        torch.cuda.synchronize()
    m1.check()

    # ---- device-resident arm
    import gc
    gc.collect()
    gc.disable()          # a generation-2 collection inside a step showed up as one 24-34 ms step among 16.4 ms ones
    barrier()
    lib.ssv_launch_count(1)
    step_ms, parts = [], []
    clk = sampler
    clk.rows.clear()                                 # keep only the samples taken during the timed region
    try:
        for _ in range(args.steps):
            flush.fill_(1)
            e, lin = device_step()
            torch.cuda.synchronize()
            step_ms.append(e[0].elapsed_time(e[3]))
            parts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
        barrier()
        launches = int(lib.ssv_launch_count(0))
        m1.check()
        total_ms = max_over_ranks(sum(step_ms))
        # ---- end-to-end arm: host buffers in, host buffers out, through ssv_synthesize_host
        # (a) one batch at a time, fully synchronous per step
        e2e_s = []
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            syn.synthesize_host(ids_np, spk_np, T)
            e2e_s.append(time.perf_counter() - t0)
        barrier()
        e2e_sync_total = max_over_ranks(sum(e2e_s))
        # (b) the corpus loop as the driver runs it: submit / wait, two batches in flight, so the D2H of step i
        #     overlaps TextEnc / decode of step i + 1 (every step still copies its inputs in and its result out
        #     inside the timed region; the L2 flush between steps is dropped: a 114 MB result per step streams
        #     through L2 anyway)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pend = None
        for _ in range(args.steps):
            cur = syn.submit(ids_np, spk_np, T)
            if pend is not None:
                syn.collect(pend)
            pend = cur
        syn.collect(pend)
        e2e_pipe = time.perf_counter() - t0
        barrier()
        e2e_total = max_over_ranks(e2e_pipe)
        d2h_bytes, h2d_bytes = int(syn.d2h_bytes), int(syn.h2d_bytes)
        e2e_fp32_total = None
        if lin_dtype != "fp32":                              # the same loop with the reference's fp32 host result
            syn32 = Synthesizer(m1, m2, ssrn_precision=args.ssrn_precision, lin_dtype="fp32")
            for _ in range(2):
                syn32.synthesize_host(ids_np, spk_np, T)
            barrier()
            t0 = time.perf_counter()
            pend = None
            for _ in range(args.steps):
                cur = syn32.submit(ids_np, spk_np, T)
                if pend is not None:
                    syn32.collect(pend)
                pend = cur
            syn32.collect(pend)
            e2e_fp32_total = max_over_ranks(time.perf_counter() - t0)
            d2h_fp32 = int(syn32.d2h_bytes)
            for _ in range(2):
                syn.synthesize_host(ids_np, spk_np, T)       # back to the headline's host buffers
            barrier()
    finally:
        sampler.__exit__(None, None, None)
    clocks = clk.summary()
    gc.enable()

    frames_total = world * B * T * args.steps
    value = frames_total / (total_ms * 1e-3)
    e2e_value = frames_total / e2e_total
    te_ms, dec_ms, ssrn_ms = (statistics.mean(p[i] for p in parts) for i in range(3))

    # ---- secondary measurements (outside the headline's timed region)
    extra = {}
    if B != 64 and B > 64:
        # the round-1 / BASELINE config 3 batch (64 utterances per step), device-resident, for continuity
        i64, s64 = ids_d[:64].contiguous(), spk_d[:64].contiguous()
        for _ in range(3):
            flush.fill_(1)
            device_step(ids_x=i64, spk_x=s64)
        torch.cuda.synchronize()
        ms64, parts64 = [], []
        for _ in range(10):
            flush.fill_(1)
            e, lin64 = device_step(ids_x=i64, spk_x=s64)
            torch.cuda.synchronize()
            ms64.append(e[0].elapsed_time(e[3]))
            parts64.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
        del lin64
        d64 = statistics.mean(p[1] for p in parts64)
        b64_bytes = T * (DECODE_WEIGHT_BYTES + 64 * DECODE_STATE_BYTES_PER_UTT)
        extra["batch64"] = {
            "what": "the same step with 64 utterances (round-1 bench shape; BASELINE config 3 decode batch), one rank, device-resident",
            "value": 64 * T / (statistics.mean(ms64) * 1e-3), "ms_per_step": statistics.mean(ms64),
            "phases_ms": {"text_encoder+begin": statistics.mean(p[0] for p in parts64), "decode": d64,
                          "ssrn": statistics.mean(p[2] for p in parts64)},
            "decode_us_per_frame": 1e3 * d64 / T,
            "decode_hbm_roofline_frac": b64_bytes / (d64 * 1e-3) / 1e9 / measured_peaks()["hbm"],
            "decode_fma_roofline_frac": 64 * T * DECODE_FLOP_PER_FRAME / (d64 * 1e-3) / 1e12 / FP32_PEAK_TFLOPS}
        for _ in range(2):
            flush.fill_(1)
            device_step()                       # back to the headline shape
        torch.cuda.synchronize()
    other = "fp32" if args.ssrn_precision == "bf16" else "bf16"
    m2.precision = other
    device_step()
    torch.cuda.synchronize()
    Ysrc = m1._state["Y"]
    for _ in range(2):
        m2(Ysrc)
    ea, eb = ev(), ev()
    flush.fill_(1)
    ea.record(); m2(Ysrc); eb.record(); torch.cuda.synchronize()
    extra[f"ssrn_{other}_ms"] = ea.elapsed_time(eb)
    m2.precision = args.ssrn_precision
    # BASELINE config 3, batch 1: latency-bound decode
    K1, V1 = m1.encode_text(ids_d[:1])
    best1 = None
    for _ in range(3):
        dec1 = m1._begin(K1, V1, spk_d[:1], T)
        ea, eb = ev(), ev()
        ea.record(); _lib.check(lib.ssv_decoder_run(dec1, T, _lib.current_stream_ptr())); eb.record(); torch.cuda.synchronize()
        ms = ea.elapsed_time(eb)
        best1 = ms if best1 is None else min(best1, ms)
    m1.check()
    # BASELINE config 1 shape end to end (one utterance, host buffers in and out): TextEnc + decode + SSRN
    for _ in range(2):
        syn.synthesize_host(ids_np[:1], spk_np[:1], T)
    t0 = time.perf_counter()
    for _ in range(5):
        syn.synthesize_host(ids_np[:1], spk_np[:1], T)
    extra["synthesize_batch1_e2e_ms"] = 1e3 * (time.perf_counter() - t0) / 5
    for _ in range(2):
        syn.synthesize_host(ids_np, spk_np, T)          # back to the headline shape (re-creates the staging buffers)
    extra["decode_batch1"] = {"ms": best1, "us_per_frame": 1e3 * best1 / T, "frames_per_s": T / (best1 * 1e-3),
                              "hbm_roofline_frac": T * (DECODE_WEIGHT_BYTES + DECODE_STATE_BYTES_PER_UTT) / (best1 * 1e-3) / 1e9 / measured_peaks()["hbm"]}
    # BASELINE config 5 (B = 32 global, N = 64, T = 217): the adversarial training step of Text2Mel, outside the headline's
    # timed region.  FP32 (highway convs forward / backward in the library's kernels, the small layers in torch ops).
    # Under torchrun the global batch of 32 is sharded over the ranks (strong scaling) and the generator's gradients
    # (24,073,584 fp32 elements) are all-reduced over NCCL from gradient hooks, overlapped with the backward pass; the
    # discriminator's 119,233 after its two backward passes.  Times are the max over ranks.
    try:
        from spoofsv_b200 import train as TR
        m1.train()
        Bg, Nt = 32, 64
        sl = TR.shard_batch(Bg, world, rank)
        gen = torch.Generator(device="cuda").manual_seed(1234)           # the same global batch on every rank
        ids_t = torch.randint(2, 34, (Bg, 1, Nt), device="cuda", generator=gen)[sl]
        spk_g = spk_d[:Bg] if spk_d.shape[0] >= Bg else spk_d[:1].expand(Bg, -1, -1).contiguous()
        spk_t = spk_g[sl].contiguous()
        mel_t = torch.rand((Bg, 80, T), device="cuda", generator=gen)[sl].contiguous()
        tgt_t = torch.rand((Bg, 80, T), device="cuda", generator=gen)[sl].contiguous()
        mel_g = (torch.rand((Bg, 80, T), device="cuda", generator=gen) * 0.9 + 0.05)[sl].contiguous()

        def timed(fn, n=3):
            for _ in range(2):
                fn()
            barrier()
            ea, eb = ev(), ev()
            ea.record()
            for _ in range(n):
                fn()
            eb.record(); torch.cuda.synchronize()
            return max_over_ranks(ea.elapsed_time(eb) / n)

        if sl.stop > sl.start:
            def train_step():
                m1.zero_grad(set_to_none=True)
                Yt, _ = m1(mel_t, ids_t, spk_t)
                (Yt - tgt_t).abs().mean().backward()

            if world == 1:
                extra["config5_generator_fwd_bwd_ms"] = timed(train_step)

            class _NoStep:        # the optimizer step is left out so that the benchmark model keeps its weights
                def __init__(self, params): self.params = list(params)
                def zero_grad(self, set_to_none=True):
                    for p in self.params: p.grad = None
                def step(self): pass

            torch.manual_seed(4321)                                     # identical discriminator on every rank
            disc = TR.melDisc(80, 128).cuda().train()
            gaw = TR.guided_attention_mat(186, 325, device="cuda")
            og, od = _NoStep(m1.parameters()), _NoStep(disc.parameters())
            sw = TR.shard_weight(Bg, world, rank)
            red = TR.OverlappedGradReducer(m1.parameters()) if world > 1 else None
            g_ms = timed(lambda: TR.generator_step(m1, disc, og, mel_g, ids_t, spk_t, gaw, {"LAMBDA": 10}, shard_weight=sw, reducer=red))
            d_ms = timed(lambda: TR.discriminator_step(m1, disc, od, mel_g, ids_t, spk_t, {"LAMBDA": 10}, shard_weight=sw))
            if red is not None:
                g_post_ms = timed(lambda: TR.generator_step(m1, disc, og, mel_g, ids_t, spk_t, gaw, {"LAMBDA": 10}, shard_weight=sw))
                red.close()
                extra["config5_g_iteration_allreduce_after_backward_ms"] = g_post_ms
            extra["config5_g_iteration_ms"] = g_ms
            extra["config5_d_iteration_ms"] = d_ms
            if world == 1:
                # the same two iterations captured in CUDA graphs and replayed (fixed shapes): issued kernel by kernel
                # the host sets the pace of both (~800 launches / several hundred small autograd ops)
                try:
                    m1.layerwise_train_forward = True
                    gg = TR.GraphedIteration(lambda mel, i, sp: TR.generator_body(m1, disc, og, mel, i, sp, gaw, {"LAMBDA": 10}),
                                             (mel_g, ids_t, spk_t), parameters=list(m1.parameters()) + list(disc.parameters()))
                    gd = TR.GraphedIteration(lambda mel, i, sp: TR.discriminator_body(m1, disc, od, mel, i, sp, {"LAMBDA": 10}),
                                             (mel_g, ids_t, spk_t), parameters=list(m1.parameters()) + list(disc.parameters()))
                    extra["config5_g_iteration_graph_ms"] = timed(lambda: gg(mel_g, ids_t, spk_t))
                    extra["config5_d_iteration_graph_ms"] = timed(lambda: gd(mel_g, ids_t, spk_t))
                    del gg, gd
                except Exception as exc:
                    extra["config5_graph_error"] = f"{type(exc).__name__}: {exc}"
                finally:
                    m1.layerwise_train_forward = False
            extra["config5"] = {"global_batch": Bg, "per_rank_batch": sl.stop - sl.start, "text_len": Nt, "frames": T, "dtype": "fp32",
                                "allreduce": "NCCL, 8 MB buckets launched from gradient hooks during backward" if world > 1 else "none (1 rank)",
                                "g_allreduce_elements": 24_073_584 if world > 1 else 0, "d_allreduce_elements": 119_233 if world > 1 else 0}
    except Exception as exc:           # context only: never lose the headline line over it
        extra["config5_error"] = f"{type(exc).__name__}: {exc}"
    finally:
        m1.eval()
        m1.zero_grad(set_to_none=True)

    peaks = measured_peaks()
    dec_bytes = T * (DECODE_WEIGHT_BYTES + B * DECODE_STATE_BYTES_PER_UTT)
    dec_gbs = dec_bytes / (dec_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text())          # ncu --set full captures of one decode launch, keyed by batch size
        traffic = tj.get("by_batch", {}).get(str(B), {}).get("decode_kernel_dram_bytes_per_launch")
        if traffic is None and B == 64:
            traffic = tj.get("decode_kernel_dram_bytes_per_launch")
    ssrn_tflops = B * T * SSRN_FLOP_PER_FRAME / (ssrn_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("fp32 (TextEnc: tcgen05 kind::tf32 with 3xTF32 split operands, FP32-accurate; decode: FFMA)"
                  + (" + fp32 SSRN (3xTF32, tcgen05)" if args.ssrn_precision == "fp32" else " + bf16 SSRN (tcgen05, fp32 accumulate)")),
        "data": "synthetic",
        "config": config_dict(args, B),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e2e_total / args.steps,
                "d2h_dtype": lin_dtype + " linear spectrogram (B, 513, 4T) + int64 alignment trajectory",
                **({"fp32_out_value": frames_total / e2e_fp32_total, "fp32_out_ms_per_step": 1e3 * e2e_fp32_total / args.steps,
                    "fp32_out_d2h_bytes_per_step": d2h_fp32} if e2e_fp32_total else {}),
                "api": "ssv_synthesize_host_submit / _wait (C ABI, pinned host buffers, two batches in flight)",
                "sync_value": frames_total / e2e_sync_total, "sync_ms_per_step": 1e3 * e2e_sync_total / args.steps,
                "sync_api": "ssv_synthesize_host (one batch at a time)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "phases_ms": {"text_encoder+begin": te_ms, "decode": dec_ms, "ssrn": ssrn_ms},
        "step_ms": [round(x, 3) for x in step_ms], "e2e_sync_step_ms": [round(1e3 * x, 3) for x in e2e_s],
        "roofline": {"kernel": "decode_ws_kernel (weight-stationary pipelined incremental Text2Mel decode)", "bound": "hbm",
                     "achieved": dec_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": dec_gbs / peaks["hbm"],
                     "traffic": traffic, "peak_source": peaks["source"],
                     "algorithmic_bytes_per_launch": dec_bytes, "us_per_frame": 1e3 * dec_ms / T,
                     "frac_at_batch64": extra.get("batch64", {}).get("decode_hbm_roofline_frac"),
                     "note": "SURVEY 8d's decode model (27.29 MB of weights once per frame + 56 KB of state per utterance) "
                             "against the measured HBM copy bandwidth; the weights never leave the SMs, so this fraction "
                             "falls with the batch by construction (frac_at_batch64: BASELINE config 3's batch, same run); "
                             "from ~32 utterances on the binding roof is the FP32 pipe: roofline_fma"},
        "roofline_fma": {"kernel": "decode_ws_kernel", "bound": "fp32 FMA", "achieved": B * T * DECODE_FLOP_PER_FRAME / (dec_ms * 1e-3) / 1e12,
                         "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": B * T * DECODE_FLOP_PER_FRAME / (dec_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
                         "peak_source": "148 SMs x 128 lanes x 2 x 1.965 GHz (no measured fp32 peak in MEASURED_PEAKS.json)"},
        "roofline_textenc": {"kernel": "conv_tf32x3_kernel x 14 (+ embedding gather, split, K/V transposes, decoder begin)",
                             "bound": "tensor", "unit": "TFLOP/s",
                             "achieved": 3 * B * N_TEXT * TEXTENC_FLOP_PER_CHAR / (te_ms * 1e-3) / 1e12,
                             "peak": peaks["bf16_sustained"] / 2,
                             "frac": 3 * B * N_TEXT * TEXTENC_FLOP_PER_CHAR / (te_ms * 1e-3) / 1e12 / (peaks["bf16_sustained"] / 2),
                             "useful_fp32_tflops": B * N_TEXT * TEXTENC_FLOP_PER_CHAR / (te_ms * 1e-3) / 1e12,
                             "note": "achieved counts the three TF32 MMAs per product (hi*hi, hi*lo, lo*hi); peak = half the measured "
                                     "sustained bf16 rate (kind::tf32 issues at half the kind::f16 rate); useful_fp32_tflops is the "
                                     "algorithmic FP32 rate (CUDA-core FP32 peak: 74.4)"},
        "roofline_ssrn": ({"bound": "tensor", "achieved": ssrn_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                           "frac": ssrn_tflops / peaks["bf16_sustained"], "precision": args.ssrn_precision}
                          if args.ssrn_precision == "bf16" else
                          {"bound": "tensor", "achieved": 3 * ssrn_tflops, "peak": peaks["bf16_sustained"] / 2, "unit": "TFLOP/s",
                           "frac": 3 * ssrn_tflops / (peaks["bf16_sustained"] / 2), "useful_fp32_tflops": ssrn_tflops,
                           "precision": "fp32 (3xTF32)",
                           "note": "achieved counts the three TF32 MMAs per product; peak = half the measured sustained bf16 rate"}),
        "extra": extra,
    }
    if rank == 0 and world == 1 and not args.no_eager:
        line["extra"]["torch_eager_gpu"] = torch_eager_gpu(sd1, sd2, ids_np, spk_np, T)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import build_ref
        cores = os.cpu_count() or 1
        nb = max(1, min(args.cpu_sample, B))
        got = syn.synthesize_host(ids_np, spk_np, T)["lin"][:nb]
        if not isinstance(got, np.ndarray):
            got = got.float().numpy()
        olin, cb = None, None
        if build_ref.available():
            # the unmodified reference module on the host cores, in a child process with the GPUs hidden (the module
            # binds `device` at import): one timed step of the bounded sample, its spectrogram kept for the check below
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                dump = os.path.join(td, "lin.npy")
                cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                       "--batch", str(B), "--frames", str(T), "--cpu-sample", str(nb), "--dump-lin", dump]
                env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, timeout=900)
                for ln in r.stdout.splitlines():
                    if ln.startswith("{"):
                        cb = json.loads(ln)["cpu_baseline"]
                if cb is not None and os.path.exists(dump):
                    olin = np.load(dump)
        if cb is None:
            torch.set_num_threads(cores)
            cpu_reference_step(sd1, sd2, ids_np[:nb], spk_np[:nb], 8)
            sec, o = cpu_reference_step(sd1, sd2, ids_np[:nb], spk_np[:nb], T)
            olin = o.numpy()
            cb = {"value": nb * T / sec, "unit": UNIT, "cores": cores, "kind": "port",
                  "sample": f"{nb} of the {B} utterances, all {T} frames, reference re-encoding AR loop + SSRN (oracle port, torch CPU fp32)"}
        cb["max_abs_err_vs_gpu_lin"] = float(np.abs(got - olin).max())
        cb["rel_l2_err_vs_gpu_lin"] = float(np.linalg.norm((got - olin).ravel()) / np.linalg.norm(olin.ravel()))
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
