"""End-to-end host->device->host time of one 2^24 transform vs coalescing width (development aid):
the first and last pass access pinned host memory directly, so their run length is the PCIe
request size."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
h_in = torch.randint(0, 2**28, (n * 8,), dtype=torch.int32).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
a, b = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)
d = h_in.cuda()
for T, g in ((11, 2), (11, 3), (11, 4), (11, 5), (12, 2), (12, 4), (12, 6), (10, 4)):
    ctx = Ntt256("bn254_fr", max_log_n=log_n)
    ctx.set_tuning(T, g)
    try:
        for _ in range(2): ctx.forward_host(a, log_n, out=b)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); ctx.forward_host(a, log_n, out=b); ts.append(time.perf_counter() - t0)
        for _ in range(2): ctx.forward(d, log_n)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): ctx.forward(d, log_n)
        torch.cuda.synchronize(); dev = (time.perf_counter() - t0) / 5
        print(json.dumps(dict(T=T, g=g, e2e_ms=round(min(ts) * 1e3, 3), device_ms=round(dev * 1e3, 3))), flush=True)
    except Exception as e:
        print(json.dumps(dict(T=T, g=g, error=str(e))), flush=True)
    ctx.close()
# plain copies for reference
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h_in, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t0
t0 = time.perf_counter(); h_out.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = time.perf_counter() - t0
print(json.dumps(dict(h2d_ms=round(h2d * 1e3, 3), d2h_ms=round(d2h * 1e3, 3), gbs=round(n * 32 / h2d / 1e9, 1))))
