"""Development aid: CUDA-event timing of TextEnc (B=64, N=58) and the fp32 SSRN arm."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, m2 = W.build_models(0); m1, m2 = m1.cuda(), m2.cuda()
ids = W.synthetic_text(64, 58, seed=11).cuda()
mel = torch.rand((64, 80, 217), device="cuda")
def t(fn, n=5):
    for _ in range(2): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"TextEnc B=64 N=58: {t(lambda: m1.encode_text(ids)):.3f} ms")
print(f"TextEnc B=1  N=58: {t(lambda: m1.encode_text(ids[:1])):.3f} ms")
print(f"SSRN fp32 B=64: {t(lambda: m2(mel), 3):.3f} ms")
