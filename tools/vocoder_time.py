"""Development aid: time the batched waveform stage at the corpus shape (64 utterances x 1304 frames)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spoofsv_b200 import vocoder as G
cfg = {"NORM_POWER": {"ANALYSIS": 0.6, "RECONSTRUCTION": 1.3}, "STFT": {"FFT_LENGTH": 1024, "HOP_LENGTH": 256},
       "PREEMPH": 0.97, "SAMPLING_RATE": 22050, "LOG_FEATURE": False}
lin = torch.rand((64, 513, 1304), device="cuda") * 0.9 + 0.05
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w = G.postprocess(lin, cfg)
    torch.cuda.synchronize(); print(f"postprocess 64 x 1304 frames: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
S = (lin / lin.amax(dim=(1, 2), keepdim=True)).pow(1.3 / 0.6)
for fused in (True, False):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = G.griffin_lim(S, 64, fused=fused)
        torch.cuda.synchronize(); print(f"griffin_lim fused={fused}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
