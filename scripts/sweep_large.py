"""2^26 / 2^28 on one GPU: tile size and streamed-first-pass-twiddle limit (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt256


def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for log_n in (26, 28):
    d = torch.randint(0, 2**28, ((1 << log_n) * 8,), dtype=torch.int32, device="cuda")
    for T, g in ((11, 2), (12, 2), (12, 0)):
        for limit_gib in (1, 16):
            ctx = Ntt256("bn254_fr", max_log_n=28)
            ctx.set_tuning(T, g)
            ctx.set_stream_twiddle_limit(limit_gib << 30)
            ms = timeit(lambda: ctx.forward(d, log_n))
            print(json.dumps(dict(log_n=log_n, T=T, g=g, stream_limit_gib=limit_gib, fwd_ms=round(ms, 3),
                                  gelem=round((1 << log_n) / ms / 1e6, 3))), flush=True)
            ctx.close()
            torch.cuda.empty_cache()
    del d
