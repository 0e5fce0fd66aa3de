"""Forward/inverse timings over sizes with the default (auto) tiling (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ctx = Ntt256("bn254_fr", max_log_n=int(sys.argv[1]) if len(sys.argv) > 1 else 24)
rng = np.random.default_rng(1)
for log_n, batch in ((10, 1), (12, 1), (14, 1), (16, 1), (18, 1), (20, 1), (22, 1), (24, 1), (16, 256), (12, 4096), (26, 1), (28, 1)):
    if log_n > ctx.max_log_n: continue
    n = batch << log_n
    x = rng.integers(0, 2**31, size=(n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
    d = torch.from_numpy(x).cuda().view(-1); del x
    f = timeit(lambda: ctx.forward(d, log_n, batch)); i = timeit(lambda: ctx.inverse(d, log_n, batch))
    print(json.dumps(dict(log_n=log_n, batch=batch, fwd_ms=round(f, 4), inv_ms=round(i, 4), fwd_gelem=round(n / f / 1e6, 3), inv_gelem=round(n / i / 1e6, 3))), flush=True)
    del d
