"""Tile x run-length sweep of the 256-bit transforms (eight elements per thread) per call size: development aid
for the defaults in capi.cu (tile_bits 11, coalesce_bits 2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt256

def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

field = sys.argv[1] if len(sys.argv) > 1 else "bn254_fr"
for log_n, batch, reps in ((24, 1, 10), (22, 1, 20), (21, 1, 30), (20, 1, 30), (19, 1, 30), (18, 1, 50), (16, 256, 10), (16, 16, 30), (14, 1024, 10), (12, 4096, 10)):
    d = torch.randint(0, 1 << 28, ((batch << log_n) * 8,), dtype=torch.int32, device="cuda")
    row = {"field": field, "log_n": log_n, "batch": batch}
    ctx = Ntt256(field, max_log_n=24)
    row["auto"] = round(timed(lambda: ctx.forward(d, log_n, batch=batch), reps), 5)
    ctx.close()
    for T in (10, 11, 12):
        for g in (0, 1, 2, 3):
            if T > log_n: continue
            ctx = Ntt256(field, max_log_n=24)
            try:
                ctx.set_tuning(T, g)
                row[f"T{T}g{g}"] = round(timed(lambda: ctx.forward(d, log_n, batch=batch), reps), 5)
            except Exception:
                row[f"T{T}g{g}"] = "err"
            ctx.close()
    best = min((v, k) for k, v in row.items() if isinstance(v, float))
    row["best"] = best[1]
    print(json.dumps(row), flush=True)
