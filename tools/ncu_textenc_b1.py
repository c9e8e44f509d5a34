import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, _ = W.build_models(0); m1 = m1.cuda()
ids = W.synthetic_text(1, 58, seed=11).cuda()
for _ in range(3):
    K, V = m1.encode_text(ids)
torch.cuda.synchronize()
print("done")
