"""Development aid: per-phase cycle breakdown of the decode kernel (SSV_DECODE_PROF=1)."""
import os, sys, time
os.environ["SSV_DECODE_PROF"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
m1, _ = W.build_models(0); m1 = m1.cuda()
names, emb, _ = W.load_fixtures()
for B in [int(a) for a in sys.argv[1:]] or [1, 64]:
    ids = W.synthetic_text(B, 58, seed=11).cuda(); spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None].cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        m1.synthesize(ids, spk, 217); torch.cuda.synchronize()
        print(f"B={B} wall {time.time()-t0:.4f}s", flush=True)
