"""Development aid: does a concurrent device->host copy (the previous batch's spectrogram) slow the decode launch?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from spoofsv_b200 import _lib
m1, _ = W.build_models(0); m1 = m1.cuda()
names, emb, _ = W.load_fixtures()
lib = _lib.load()
T, B = 217, 64
ids = W.synthetic_text(B, 58, seed=11).cuda()
spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None].cuda()
K, V = m1.encode_text(ids)
src = torch.empty(57 << 20, dtype=torch.uint8, device="cuda")
dst = torch.empty(57 << 20, dtype=torch.uint8, pin_memory=True)
side = torch.cuda.Stream()
for mode in ("alone", "with D2H 57 MB", "alone", "with D2H 57 MB", "with D2H 114 MB"):
    ts = []
    for rep in range(5):
        dec = m1._begin(K, V, spk, T)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode != "alone":
            with torch.cuda.stream(side):
                dst.copy_(src, non_blocking=True)
                if "114" in mode:
                    dst.copy_(src, non_blocking=True)
        _lib.check(lib.ssv_decoder_run(dec, T, _lib.current_stream_ptr())); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{mode:18s} decode {min(ts):7.3f} ms (median {sorted(ts)[2]:7.3f})", flush=True)
