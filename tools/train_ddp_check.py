"""Development aid (torchrun, one process per GPU): the data-parallel generator iteration equals the single-process one.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_ddp_check.py
"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from oracle import weights as W
from spoofsv_b200 import train as TR

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
B, N, T = 32, 64, 217
m1, _ = W.build_models(0); m1 = m1.cuda().train()
torch.manual_seed(5)
disc = TR.melDisc(80, 128).cuda().eval()
names, emb, _ = W.load_fixtures()
ids = W.synthetic_text(B, N, seed=3).cuda()
spk = torch.from_numpy(emb[:B].copy())[:, :, None].cuda()
mel_gt = (torch.rand((B, 80, T), generator=torch.Generator().manual_seed(1)) * 0.9 + 0.05).cuda()
gaw = TR.guided_attention_mat(186, 325, device="cuda")
cfg = {"LAMBDA": 10}
sl = TR.shard_batch(B, world, rank)
opt = torch.optim.SGD(m1.parameters(), lr=0.0)

def step(sel, group_on):
    if group_on:
        return TR.generator_step(m1, disc, opt, mel_gt[sel], ids[sel], spk[sel], gaw, cfg)
    # single-process reference: bypass the collectives by running on the full batch in every rank
    saved = dist.is_initialized
    dist.is_initialized = lambda: False
    try:
        return TR.generator_step(m1, disc, opt, mel_gt[sel], ids[sel], spk[sel], gaw, cfg)
    finally:
        dist.is_initialized = saved

full = step(slice(0, B), False)
g_full = {n: p.grad.detach().clone() for n, p in m1.named_parameters()}
for _ in range(2):
    par = step(sl, True)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(5):
    par = step(sl, True)
torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 5
errs = sorted(((float((p.grad - g_full[n]).abs().max()) / max(float(g_full[n].abs().max()), 1e-9), n, float(g_full[n].abs().max()))
               for n, p in m1.named_parameters()), reverse=True)
worst = errs[0][0]
if rank == 0:
    for e, n, mag in errs[:4]:
        print(f"  {n}: rel {e:.2e} (largest gradient {mag:.2e})")
    print("  median rel", errs[len(errs) // 2][0])
if rank == 0:
    print(f"world {world}: loss terms full {full} vs parallel {par}")
    print(f"worst relative gradient difference (allreduced shards vs full batch): {worst:.2e}")
    print(f"data-parallel G iteration: {1e3 * dt:.2f} ms for the global batch of {B}")
dist.destroy_process_group()
