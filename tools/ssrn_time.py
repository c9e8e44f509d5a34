"""Development aid: CUDA-event timing of the SSRN arms (B x 217 mel frames -> B x 513 x 868)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, m2 = W.build_models(0); m2 = m2.cuda()
def t(fn, n=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in sys.argv[1:] or ["bf16"]:
    m2.precision = prec
    for B in (32, 64):
        mel = torch.rand((B, 80, 217), device="cuda")
        ms = t(lambda: m2(mel), 10 if prec == "bf16" else 3)
        print(f"SSRN {prec} B={B}: {ms:.3f} ms  {B * 217 * 46.469144e6 / ms / 1e9:.0f} TFLOP/s", flush=True)
