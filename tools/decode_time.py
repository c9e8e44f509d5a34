"""Development aid: CUDA-event timing of ssv_decoder_run for several batch sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from spoofsv_b200 import _lib
m1, _ = W.build_models(0); m1 = m1.cuda()
names, emb, _ = W.load_fixtures()
lib = _lib.load()
T = 217
# arguments: batch sizes; "B:R:W[:F]" forces the front-end shape (ssv_decoder_set_plan)
for arg in sys.argv[1:] or ["1", "64"]:
    B, R, Wp, F = ([int(x) for x in arg.split(":")] + [0, 0, 0])[:4]
    m1.decode_plan = (R, Wp, F) if (R or Wp or F) else None
    ids = W.synthetic_text(B, 58, seed=11).cuda()
    spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None].cuda()
    K, V = m1.encode_text(ids)
    best = 1e9
    for rep in range(4):
        dec = m1._begin(K, V, spk, T)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.ssv_decoder_run(dec, T, _lib.current_stream_ptr())); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    m1.check()
    print(f"B={B:4d} plan R={R} W={Wp} F={F} decode {best:8.3f} ms  {1e3 * best / T:7.1f} us/frame  {B * T / best * 1e3:12.0f} frames/s", flush=True)
