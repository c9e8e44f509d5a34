"""Development aid: CUDA-event timing of the fused Griffin-Lim at the corpus shape (64 x 513 x 1304, 64 iterations)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spoofsv_b200 import vocoder as G
B, T = 64, 1304
S = torch.rand((B, 513, T), device="cuda")
ph = torch.rand((B, 513, T), device="cuda") * 6.2831853
a0 = torch.polar(torch.ones_like(ph), ph)
for n_iter in (64, 0):
    for rep in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        y = G.griffin_lim(S, n_iter, angles0=a0)
        e1.record(); torch.cuda.synchronize()
        print(f"n_iter={n_iter} fused: {e0.elapsed_time(e1):.2f} ms", flush=True)
