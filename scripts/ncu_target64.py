"""Minimal single-GPU program for ncu: Goldilocks forward NTTs through the C ABI (argv: log_n batch iters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api64 import Ntt64
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = Ntt64("goldilocks", max_log_n=max(log_n, 16))
d = torch.randint(0, 1 << 40, (batch << log_n,), dtype=torch.int64, device="cuda")
for _ in range(iters):
    ctx.forward(d, log_n, batch=batch)
torch.cuda.synchronize()
print("done", ctx.launch_count)
