"""Development aid: TextEnc at the bench shape (run under ncu -k regex:conv_tf32x3, or conv_f32 with SSV_TE_PREC=fp32-ffma)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, _ = W.build_models(0); m1 = m1.cuda(); m1.precision = os.environ.get("SSV_TE_PREC", "fp32")
ids = W.synthetic_text(64, 58, seed=11).cuda()
for _ in range(2):
    K, V = m1.encode_text(ids)
torch.cuda.synchronize()
print("done")
