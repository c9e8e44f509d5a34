"""Development aid: one generator forward + backward at the config-5 shape (run under ncu for a launch list)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
B, N, T = 32, 64, 217
m1, _ = W.build_models(0); m1 = m1.cuda().train()
names, emb, _ = W.load_fixtures()
ids = W.synthetic_text(B, N, seed=3).cuda()
spk = torch.from_numpy(emb[:B].copy())[:, :, None].cuda()
mel = torch.rand((B, 80, T), device="cuda"); tgt = torch.rand((B, 80, T), device="cuda")
for _ in range(2):
    m1.zero_grad(set_to_none=True)
    Y, A = m1(mel, ids, spk)
    ((Y - tgt).abs().mean()).backward()
torch.cuda.synchronize()
print("done")
