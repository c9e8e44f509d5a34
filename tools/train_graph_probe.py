"""Development aid: capture the G and D iterations (config-5 shape) in CUDA graphs, compare one replay with the eager
iteration from the same state, and time both."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from spoofsv_b200 import train as TR
B, N, T = 32, 64, 217
m1, _ = W.build_models(0); m1 = m1.cuda().train()
names, emb, _ = W.load_fixtures()
ids = W.synthetic_text(B, N, seed=3).cuda()
spk = torch.from_numpy(emb[:B].copy())[:, :, None].cuda()
mel_gt = torch.rand((B, 80, T), device="cuda") * 0.9 + 0.05
coeff = torch.rand(B, device="cuda")
disc = TR.melDisc(80, 128).cuda().train()
if '--dropout' not in sys.argv:
    for m_ in disc.modules():              # the comparison needs a deterministic discriminator
        if isinstance(m_, torch.nn.Dropout): m_.p = 0.0
gaw = TR.guided_attention_mat(186, 325, device="cuda")
cfg = {"LAMBDA": 10}
opt_g = torch.optim.Adam(m1.parameters(), 2e-4, (0.5, 0.9), 1e-6, capturable=True)
opt_d = torch.optim.Adam(disc.parameters(), 2e-4, (0.5, 0.9), 1e-6, capturable=True)

def snapshot(mods, opts):
    out = [p.detach().clone() for m in mods for p in m.parameters()]
    for o in opts:
        for st in o.state.values():
            out += [v.clone() for v in st.values() if torch.is_tensor(v)]
    return out
def restore(mods, opts, snap):
    it = iter(snap)
    with torch.no_grad():
        for m in mods:
            for p in m.parameters(): p.copy_(next(it))
        for o in opts:
            for st in o.state.values():
                for v in st.values():
                    if torch.is_tensor(v): v.copy_(next(it))

def timeit(fn, n=10):
    for _ in range(2): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

m1.layerwise_train_forward = True
for name, body, opt, mods, inputs, to_dict in (
        ("G", lambda mel, i, s: TR.generator_body(m1, disc, opt_g, mel, i, s, gaw, cfg), opt_g, (m1,), (mel_gt, ids, spk), TR.generator_terms),
        ("D", lambda mel, i, s, c: TR.discriminator_body(m1, disc, opt_d, mel, i, s, cfg, coeff=c), opt_d, (disc,), (mel_gt, ids, spk, coeff), TR.wgan_terms)):
    try:
        g = TR.GraphedIteration(body, inputs, parameters=list(m1.parameters()) + list(disc.parameters()))
    except Exception as exc:
        print(f"{name}: capture failed: {type(exc).__name__}: {exc}", flush=True)
        continue
    snap = snapshot(mods, (opt,))
    out_g = to_dict(g(*inputs)); params_g = [p.detach().clone() for m in mods for p in m.parameters()]
    restore(mods, (opt,), snap)
    out_e = to_dict(body(*inputs)); params_e = [p.detach().clone() for m in mods for p in m.parameters()]
    dmax = max(float((a - b).abs().max()) for a, b in zip(params_g, params_e))
    print(f"{name}: graph {out_g}\n{name}: eager {out_e}\n{name}: max |param difference| after one step {dmax:.3e}", flush=True)
    print(f"{name} iteration: graph replay {timeit(lambda: g(*inputs)):.2f} ms, eager {timeit(lambda: to_dict(body(*inputs))):.2f} ms", flush=True)
