"""Development aid: tcgen05 path vs golden, per case."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import weights as W
from spoofsv_b200.models import highwayConv
G = W.GOLDEN
def rel(a, b):
    a = a.detach().cpu().double(); b = torch.as_tensor(b).double(); return float((a - b).norm() / b.norm())
z = np.load(G / "highway_cases.npz")
for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
    hc = highwayConv(d, k, dil, bool(causal)); hc.load_state_dict(W.highway_params(d, k, 100 + i)); hc = hc.cuda(); hc.precision = "bf16"
    x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
    try:
        y = hc(x.cuda()); torch.cuda.synchronize()
        print("hc bf16", (d, k, dil, causal), "rel-L2 %.3e  maxabs %.3e" % (rel(y, z[f"y{i}"]), float((y.cpu() - torch.as_tensor(z[f"y{i}"])).abs().max())), flush=True)
    except Exception as e:
        print("hc bf16", (d, k, dil, causal), "FAILED", e, flush=True); break
m1, m2 = W.build_models(7, init="kaiming", ln_jitter=True); m2 = m2.cuda()
zz = np.load(G / "ssrn_seed7.npz")
m2.precision = "bf16"
try:
    lin = m2(torch.from_numpy(zz["mel"]).cuda()); torch.cuda.synchronize()
    print("ssrn bf16 rel-L2 %.3e maxabs %.3e" % (rel(lin, zz["lin"]), float((lin.cpu() - torch.as_tensor(zz["lin"])).abs().max())), flush=True)
    mel = torch.rand((32, 80, 217), device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time(); lin = m2(mel); torch.cuda.synchronize()
        print("ssrn 32x217 bf16: %.4fs" % (time.time() - t0), flush=True)
    mel = torch.rand((64, 80, 217), device="cuda")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time(); lin = m2(mel); torch.cuda.synchronize()
        print("ssrn 64x217 bf16: %.4fs" % (time.time() - t0), flush=True)
except Exception as e:
    print("ssrn bf16 FAILED", e, flush=True)
