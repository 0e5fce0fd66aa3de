"""Minimal single-GPU program for ncu: two 2^22 forward transforms over BLS12-377 Fq (384-bit containers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt384
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = Ntt384("bls12_377_fq", max_log_n=log_n)
d = torch.randint(0, 2**31 - 1, ((1 << log_n), 12), dtype=torch.int32, device="cuda")
d[:, 11] &= 0x00FFFFFF
d = d.view(-1)
for _ in range(2):
    ctx.forward(d, log_n)
torch.cuda.synchronize()
print("done", ctx.launch_count)
