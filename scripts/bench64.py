"""Device time of the 64-bit transforms (Goldilocks and the reference's modulus) over sizes and batches, with the
two rooflines that bound them: HBM (16 B per element per pass, measured copy peak) and the integer pipes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api64 import Ntt64, plan_describe64

def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

peak = 6531.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
tunings = [None]
if len(sys.argv) > 1 and sys.argv[1] == "sweep":
    tunings = [None, (12, 4), (12, 3), (13, 3), (14, 2), (14, 3), (11, 4), (10, 4), (12, 2)]
for field in ("goldilocks", "ref29"):
    ctx = Ntt64(field, max_log_n=26 if field == "goldilocks" else 24)
    for tun in tunings:
        if tun is not None:
            ctx.set_tuning(*tun)
        for log_n, batch, reps in ((24, 1, 20), (26, 1, 5), (22, 1, 20), (20, 1, 50), (16, 256, 20), (16, 1, 100), (12, 4096, 20), (12, 1, 100)):
            if log_n > ctx.max_log_n: continue
            n = batch << log_n
            d = torch.randint(0, 1 << 28, (n,), dtype=torch.int64, device="cuda")
            ms = timed(lambda: ctx.forward(d, log_n, batch=batch), reps)
            d = torch.randint(0, 1 << 28, (n,), dtype=torch.int64, device="cuda")
            msi = timed(lambda: ctx.inverse(d, log_n, batch=batch), reps)
            passes = None
            if tun is None:
                passes = len(plan_describe64(log_n, batch, ctx.max_log_n)["passes"])
            rec = {"field": field, "tuning": tun, "log_n": log_n, "batch": batch, "fwd_ms": round(ms, 4), "inv_ms": round(msi, 4),
                   "G_elements_s": round(n / ms / 1e6, 3), "G_butterflies_s": round(n / 2 * log_n / ms / 1e6, 2)}
            if passes:
                rec["passes"] = passes
                rec["hbm_GBs_algorithmic"] = round(16 * n * passes / ms / 1e6, 1)
                rec["hbm_frac"] = round(16 * n * passes / ms / 1e6 / peak, 3)
            print(json.dumps(rec), flush=True)
    ctx.close()
