"""Tile / coalescing sweep of the 64-bit transforms per call size (development aid for choose_shape in capi64.cu)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api64 import Ntt64

def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

sizes = ((26, 1, 5), (24, 1, 20), (22, 1, 30), (21, 1, 30), (20, 1, 50), (19, 1, 50), (18, 1, 50), (17, 1, 50), (16, 1, 100), (14, 1, 100),
         (16, 256, 20), (16, 16, 50), (12, 4096, 20), (12, 256, 50), (20, 16, 20))
for log_n, batch, reps in sizes:
    n = batch << log_n
    d = torch.randint(0, 1 << 28, (n,), dtype=torch.int64, device="cuda")
    row = {"log_n": log_n, "batch": batch}
    ctx = Ntt64("goldilocks", max_log_n=26)
    row["auto"] = round(timed(lambda: ctx.forward(d, log_n, batch=batch), reps), 5)
    ctx.close()
    for T, g in ((12, 4), (12, 3), (12, 2), (13, 4), (13, 3), (13, 2), (14, 3), (14, 2), (11, 3), (10, 2), (10, 0), (8, 0)):
        if T > log_n + 0: 
            continue
        ctx = Ntt64("goldilocks", max_log_n=26)
        try:
            ctx.set_tuning(T, g)
            row[f"T{T}g{g}"] = round(timed(lambda: ctx.forward(d, log_n, batch=batch), reps), 5)
        except Exception as e:
            row[f"T{T}g{g}"] = "err"
        ctx.close()
    best = min((v, k) for k, v in row.items() if isinstance(v, float))
    row["best"] = best[1]
    print(json.dumps(row), flush=True)
