"""Development aid: forward + backward of the Text2Mel generator step at the BASELINE config-5 shape
(B = 32, N = 64, T = 217) -- spoofsv_b200 kernels vs the same graph in torch ops on the same GPU."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from oracle import ttsmodel_oracle as O
B, N, T = 32, 64, 217
m1, _ = W.build_models(0); m1 = m1.cuda().train()
names, emb, _ = W.load_fixtures()
ids = W.synthetic_text(B, N, seed=3).cuda()
spk = torch.from_numpy(emb[:B].copy())[:, :, None].cuda()
mel = torch.rand((B, 80, T), device="cuda"); tgt = torch.rand((B, 80, T), device="cuda")

def step_ours():
    m1.zero_grad(set_to_none=True)
    Y, A = m1(mel, ids, spk)
    ((Y - tgt).abs().mean()).backward()

sd = {k: v.detach().clone().requires_grad_(True) for k, v in m1.state_dict().items()}
def step_torch():
    for v in sd.values(): v.grad = None
    Y, A = O.melsyn_train_forward(sd, mel, ids, spk)
    ((Y - tgt).abs().mean()).backward()

def timeit(fn, n=5):
    for _ in range(2): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    t_issue = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    print("   (host issue time %.2f ms per call)" % (1e3 * (t_issue - t0) / n), flush=True)
    return e0.elapsed_time(e1) / n, 1e3 * (time.perf_counter() - t0) / n

print("ours  (fwd+bwd) gpu %.2f ms, wall %.2f ms" % timeit(step_ours), flush=True)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
print("torch fp32      gpu %.2f ms, wall %.2f ms" % timeit(step_torch), flush=True)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
print("torch tf32      gpu %.2f ms, wall %.2f ms" % timeit(step_torch), flush=True)
with torch.no_grad():
    print("ours  fwd only (fused train fwd) gpu %.2f ms, wall %.2f ms" % timeit(lambda: m1(mel, ids, spk)), flush=True)

# ---- whole iterations (losses, backward, Adam), as train/adversarial_wasserstein_gp.py runs them
from spoofsv_b200 import train as TR
disc = TR.melDisc(80, 128).cuda().train()
gaw = TR.guided_attention_mat(186, 325, device="cuda")
cfg = {"LAMBDA": 10}
opt_g = torch.optim.Adam(m1.parameters(), 2e-4, (0.5, 0.9), 1e-6)
opt_d = torch.optim.Adam(disc.parameters(), 2e-4, (0.5, 0.9), 1e-6)
mel_gt = torch.rand((B, 80, T), device="cuda") * 0.9 + 0.05
m1.train()
print("G iteration  gpu %.2f ms, wall %.2f ms" % timeit(lambda: TR.generator_step(m1, disc, opt_g, mel_gt, ids, spk, gaw, cfg)), flush=True)
print("D iteration  gpu %.2f ms, wall %.2f ms" % timeit(lambda: TR.discriminator_step(m1, disc, opt_d, mel_gt, ids, spk, cfg)), flush=True)
