"""Development aid: synchronous vs pipelined host-buffer synthesis loop, with the completion time of every batch."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import weights as W
from spoofsv_b200.synth import Synthesizer
m1, m2 = W.build_models(0); m1, m2 = m1.cuda(), m2.cuda(); m2.precision = "bf16"
names, emb, _ = W.load_fixtures()
B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 217
ids = W.synthetic_text(B, 58, seed=11).numpy()[:, 0, :]
spk = emb[[i % len(emb) for i in range(B)]].copy()
syn = Synthesizer(m1, m2, ssrn_precision="bf16", lin_dtype=(sys.argv[2] if len(sys.argv) > 2 else "bf16"))
for _ in range(4):
    syn.synthesize_host(ids, spk, T)
torch.cuda.synchronize()
K = 12
t0 = time.perf_counter()
for _ in range(K):
    syn.synthesize_host(ids, spk, T)
print("sync  ms/step", 1e3 * (time.perf_counter() - t0) / K)
t0 = time.perf_counter(); pend = None; done = []
for i in range(K):
    cur = syn.submit(ids, spk, T)
    if pend is not None:
        syn.collect(pend); done.append(time.perf_counter())
    pend = cur
syn.collect(pend); done.append(time.perf_counter())
print("pipe  ms/step", 1e3 * (done[-1] - t0) / K)
print("batch completion intervals ms:", [round(1e3 * (b - a), 2) for a, b in zip(done, done[1:])])
