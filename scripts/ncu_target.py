"""Minimal single-GPU program for ncu: a few 2^24 forward NTTs through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = Ntt256("bn254_fr", max_log_n=log_n)
if tile: ctx.set_tuning(tile, 2)
rng = np.random.default_rng(1)
x = rng.integers(0, 2**31, size=(1 << log_n, 8), dtype=np.int64).astype(np.int32)
x[:, 7] &= 0x0FFFFFFF
d = torch.from_numpy(x).cuda()
for _ in range(iters):
    ctx.forward(d, log_n)
torch.cuda.synchronize()
print("done", ctx.launch_count)
