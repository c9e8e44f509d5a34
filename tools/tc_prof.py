import os, sys
os.environ["SSV_TC_PROF"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, m2 = W.build_models(0); m2 = m2.cuda(); m2.precision = "bf16"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mel = torch.rand((B, 80, 217), device="cuda")
for rep in range(2):
    print("---- rep", rep, file=sys.stderr, flush=True)
    m2(mel); torch.cuda.synchronize()
