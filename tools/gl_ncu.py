"""Development aid: two Griffin-Lim iterations at the corpus shape (run under ncu -k regex:gl_)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from spoofsv_b200 import vocoder as G
B, T = 64, 1304
S = torch.rand((B, 513, T), device="cuda")
ph = torch.rand((B, 513, T), device="cuda") * 6.2831853
y = G.griffin_lim(S, 2, angles0=torch.polar(torch.ones_like(ph), ph))
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
