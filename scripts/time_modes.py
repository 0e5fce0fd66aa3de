"""Device time of the BN254 transforms in both twiddle representations (Shoup pairs vs Montgomery)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256

def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ctx = Ntt256("bn254_fr", max_log_n=24)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for mode in (True, False):
    ctx.set_twiddle_mode(mode)
    for log_n, batch, reps in ((24, 1, 20), (22, 1, 20), (20, 1, 50), (16, 256, 20), (16, 1, 100), (12, 4096, 20)):
        d = torch.randint(0, 1 << 31, ((batch << log_n), 8), dtype=torch.int32, device="cuda", generator=gen)
        d[:, 7] &= 0x0FFFFFFF
        d = d.view(-1)
        ms = timed(lambda: ctx.forward(d, log_n, batch=batch), reps)
        msi = timed(lambda: ctx.inverse(d, log_n, batch=batch), reps)
        print(json.dumps({"mode": ctx.twiddle_mode, "log_n": log_n, "batch": batch, "fwd_ms": round(ms, 4), "inv_ms": round(msi, 4),
                          "G_elements_s": round((batch << log_n) / ms / 1e6, 3)}), flush=True)
