"""Device time of one workload on every pass-kernel shape: elements per thread x tile size x the dense
(three CTAs per SM) instance of the four-elements-per-thread kernel.  Development aid for the shape table
in ntt_b200/csrc/capi.cu transform()."""
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
ctx = Ntt256(field, max_log_n=24)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
sizes = ((24, 1, 10), (21, 1, 30), (16, 256, 10), (12, 4096, 10), (20, 1, 30))
if len(sys.argv) > 2 and sys.argv[2] == 'mid':
    sizes = ((18, 1, 50), (19, 1, 50), (20, 1, 50), (21, 1, 30), (22, 1, 20), (16, 4, 50), (16, 16, 50), (16, 32, 30), (12, 256, 50), (12, 1024, 30))
for log_n, batch, reps in sizes:
    d = torch.randint(0, 1 << 31, ((batch << log_n), 8), dtype=torch.int32, device="cuda", generator=gen)
    d[:, 7] &= 0x0FFFFFFF
    d = d.view(-1)
    for gb, T, dense in ((3, 11, 0), (3, 10, 0), (3, 9, 0), (2, 10, 0), (2, 10, 1), (2, 9, 1), (2, 9, 0)):
        if T > log_n: continue
        ctx.set_tuning(T, 2)
        ctx.set_elements_per_thread(gb)
        ctx.set_dense_pass4(bool(dense))
        try:
            ms = timed(lambda: ctx.forward(d, log_n, batch=batch), reps)
        except Exception as e:
            print(json.dumps({"log_n": log_n, "batch": batch, "gb": gb, "T": T, "dense": dense, "error": str(e)}), flush=True)
            continue
        print(json.dumps({"field": field, "mode": ctx.twiddle_mode, "log_n": log_n, "batch": batch, "elements_per_thread": 1 << gb, "tile_bits": T,
                          "dense": dense, "ms": round(ms, 4), "G_elements_s": round((batch << log_n) / ms / 1e6, 3)}), flush=True)
