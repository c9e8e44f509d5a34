"""Development aid: one decode launch of a -DSSV_WS_TOTALS build (SSV_B200_LIB=...), cycles per stage visit and role."""
import os, sys
os.environ["SSV_DECODE_TOTALS"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from oracle import weights as W
from spoofsv_b200 import _lib
m1, _ = W.build_models(0); m1 = m1.cuda()
names, emb, _ = W.load_fixtures()
lib = _lib.load()
T = 217
for arg in sys.argv[1:] or ["64"]:
    B, R, Wp, F = ([int(x) for x in arg.split(":")] + [0, 0, 0])[:4]
    m1.decode_plan = (R, Wp, F) if (R or Wp or F) else None
    ids = W.synthetic_text(B, 58, seed=11).cuda()
    spk = torch.from_numpy(emb[[i % len(emb) for i in range(B)]].copy())[:, :, None].cuda()
    K, V = m1.encode_text(ids)
    dec = m1._begin(K, V, spk, T)
    _lib.check(lib.ssv_decoder_run(dec, T, _lib.current_stream_ptr()))
    torch.cuda.synchronize()
