"""Sweep tile_bits / coalesce_bits for several sizes (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256

def timeit(fn, reps=7):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

ctx = Ntt256("bn254_fr", max_log_n=26)
rng = np.random.default_rng(1)
res = []
for log_n, batch in ((12, 4096), (16, 256), (18, 64), (20, 16), (20, 1), (22, 1), (24, 1), (26, 1)):
    n = batch << log_n
    x = rng.integers(0, 2**31, size=(n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
    d = torch.from_numpy(x).cuda().view(-1)
    for T in (9, 10, 11, 12):
        for g in (1, 2, 3):
            if T > log_n and T != 12: continue
            try:
                ctx.set_tuning(T, g)
                ms = timeit(lambda: ctx.forward(d, log_n, batch))
            except Exception as e:
                print("fail", log_n, batch, T, g, e); continue
            res.append(dict(log_n=log_n, batch=batch, T=T, g=g, ms=ms, gelem=n / ms / 1e6))
            print(json.dumps(res[-1]), flush=True)
    del d
