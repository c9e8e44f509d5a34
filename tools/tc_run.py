import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, m2 = W.build_models(0); m2 = m2.cuda(); m2.precision = "bf16"
mel = torch.rand((int(sys.argv[1]) if len(sys.argv) > 1 else 32, 80, 217), device="cuda")
for rep in range(3):
    m2(mel); torch.cuda.synchronize()
print("ok")
