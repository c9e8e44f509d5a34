"""FP32-accurate tensor-core arm (tcgen05 kind::tf32, split operands: csrc/conv_tc32.cu) against the CUDA-core FFMA arm and
the reference's golden vectors: how far are K / V, and do the alignment trajectories stay identical?

    python tools/split_mma_probe.py            # on a B200

Prints, for the kaiming-init ragged batch (tests/golden/small_seed7.npz) and BASELINE config 1 (cfg1_seed0.npz):
max-abs of K, V against the golden for both arms; then the B = 64, N = 58, T = 217 bench workload: both arms' K / V against
the CPU oracle, trajectory identity against the oracle's incremental loop, and the TextEnc time of both arms."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from oracle import ttsmodel_oracle as O
from oracle import weights as W

G = Path(__file__).resolve().parent.parent / "tests" / "golden"
t = lambda a: torch.from_numpy(np.asarray(a))
mx = lambda a, b: float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    for name, init in (("small_seed7", dict(seed=7, init="kaiming", ln_jitter=True)), ("cfg1_seed0", dict(seed=0))):
        z = np.load(G / f"{name}.npz")
        m1, _ = W.build_models(**init)
        m1 = m1.cuda()
        ids = t(z["textid"]).cuda()
        for prec in ("fp32", "fp32-ffma"):
            m1.precision = prec
            K, V = m1.encode_text(ids)
            line = f"{name:12s} TextEnc {prec:9s}: max|dK| {mx(K, t(z['K'])):.2e}"
            if "V" in z.files:
                line += f"  max|dV| {mx(V, t(z['V'])):.2e}"
            print(line, flush=True)
    # bench workload
    m1, _ = W.build_models(0)
    sd1 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    m1 = m1.cuda()
    names, emb, _ = W.load_fixtures()
    B, N, T = 64, 58, 217
    ids = W.synthetic_text(B, N, seed=11)
    spk = torch.from_numpy(emb[:B].copy())[:, :, None]
    with torch.no_grad():
        oK, oV = O.text_encoder(ids, sd1)
        oY, oA, otraj = O.ar_loop_incremental(sd1, ids, spk, T)
    for prec in ("fp32", "fp32-ffma"):
        m1.precision = prec
        K, V = m1.encode_text(ids.cuda())
        Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
        same = bool(np.array_equal(traj.cpu().numpy(), otraj.numpy()))
        ms = timed(lambda: m1.encode_text(ids.cuda(), check=False))
        ms1 = timed(lambda: m1.encode_text(ids[:1].cuda(), check=False))
        print(f"B=64 N=58    TextEnc {prec:9s}: max|dK| {mx(K, oK):.2e}  max|dV| {mx(V, oV):.2e}  mel max-abs {mx(Y, oY):.2e}  "
              f"att {mx(A, oA):.2e}  trajectories identical over {T} frames x {B}: {same}   {ms:.3f} ms (B=1: {ms1:.3f} ms)", flush=True)


if __name__ == "__main__":
    main()
