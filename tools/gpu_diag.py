"""Verbose per-component parity report (development aid; run on the GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from oracle import ttsmodel_oracle as O
from oracle import weights as W

G = W.GOLDEN
def err(a, b): return float((a.detach().cpu().float() - torch.as_tensor(b).float()).abs().max())

def main():
    from spoofsv_b200.models import highwayConv
    print(torch.cuda.get_device_name(0))
    z = np.load(G / "highway_cases.npz")
    for i, (d, k, dil, causal) in enumerate(z["cases"].tolist()):
        hc = highwayConv(d, k, dil, bool(causal)); hc.load_state_dict(W.highway_params(d, k, 100 + i)); hc = hc.cuda()
        x = torch.randn((2, d, 45), generator=torch.Generator().manual_seed(200 + i))
        print("hc", (d, k, dil, causal), "err %.2e" % err(hc(x.cuda()), z[f"y{i}"]), flush=True)
    m1, m2 = W.build_models(7, init="kaiming", ln_jitter=True)
    sd1 = {k: v.clone() for k, v in m1.state_dict().items()}; sd2 = {k: v.clone() for k, v in m2.state_dict().items()}
    m1, m2 = m1.cuda(), m2.cuda()
    z = np.load(G / "ssrn_seed7.npz")
    print("ssrn err %.2e" % err(m2(torch.from_numpy(z["mel"]).cuda()), z["lin"]), flush=True)
    z = np.load(G / "small_seed7.npz")
    ids, spk = torch.from_numpy(z["textid"]), torch.from_numpy(z["spk"])
    K, V = m1.encode_text(ids.cuda())
    print("textenc K %.2e V %.2e" % (err(K, z["K"]), err(V, z["V"])), flush=True)
    T = z["Y"].shape[-1]
    t0 = time.time()
    Y, A, traj, _, _ = m1.synthesize(ids.cuda(), spk.cuda(), T)
    torch.cuda.synchronize()
    print("decode small: %.3fs  Y %.2e A %.2e traj_equal %s" % (time.time() - t0, err(Y, z["Y"]), err(A, z["A"]),
          np.array_equal(traj.cpu().numpy(), z["traj"])), flush=True)
    if not np.array_equal(traj.cpu().numpy(), z["traj"]):
        print(traj.cpu().numpy().T, "\n", z["traj"].T)
    for t in range(0, T, 8):
        print("  frame", t, "Yerr %.2e" % err(Y[:, :, t], z["Y"][:, :, t]))
    m1, m2 = W.build_models(0); m1, m2 = m1.cuda(), m2.cuda()
    z = np.load(G / "cfg1_seed0.npz")
    for rep in range(2):
        t0 = time.time()
        Y, A, traj, _, _ = m1.synthesize(torch.from_numpy(z["textid"]).cuda(), torch.from_numpy(z["spk"]).cuda(), 217)
        torch.cuda.synchronize()
        print("cfg1 B=1 217 frames: %.4fs  Y %.2e A %.2e traj_equal %s" % (time.time() - t0, err(Y, z["Y"]), err(A, z["A"]),
              np.array_equal(traj.cpu().numpy(), z["traj"])), flush=True)
    names, emb, _ = W.load_fixtures()
    ids = W.synthetic_text(64, 58, seed=11).cuda(); spk = torch.from_numpy(emb[:64].copy())[:, :, None].cuda()
    for rep in range(2):
        t0 = time.time(); Y, A, traj, _, _ = m1.synthesize(ids, spk, 217); torch.cuda.synchronize()
        print("B=64 217 frames: %.4fs" % (time.time() - t0), flush=True)
    mel = torch.rand((32, 80, 217), device="cuda")
    for rep in range(2):
        t0 = time.time(); lin = m2(mel); torch.cuda.synchronize()
        print("ssrn 32x217 fp32: %.4fs" % (time.time() - t0), flush=True)

if __name__ == "__main__":
    main()
