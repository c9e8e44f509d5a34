"""Development aid: one decode launch at the bench shape (B = 64, N = 58, 217 frames) after a warm-up one
(run under ncu -k regex:decode_ws_kernel -s 1 -c 1)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from spoofsv_b200 import _lib
m1, _ = W.build_models(0); m1 = m1.cuda()
names, emb, _ = W.load_fixtures()
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ids = W.synthetic_text(B, 58, seed=11).cuda()
spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None].cuda()
K, V = m1.encode_text(ids)
for _ in range(2):
    dec = m1._begin(K, V, spk, 217)
    _lib.check(lib.ssv_decoder_run(dec, 217, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
m1.check()
print("done")
