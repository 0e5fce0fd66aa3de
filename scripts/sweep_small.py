"""Sweep (elements per thread, tile bits, coalescing bits) for small transforms; GPU time per call
from CUDA-graph replays (development aid; the winners feed the heuristic in capi.cu transform())."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256
from ntt_b200.lib import Ntt256Error


def graph_time(fn, k=10, reps=8):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 4] / k * 1e3


rng = np.random.default_rng(1)
cases = [(8, 1), (10, 1), (12, 1), (14, 1), (16, 1), (18, 1), (20, 1), (10, 64), (12, 64)]
mid = len(sys.argv) > 1 and sys.argv[1] == "mid"  # the cross-over region between the kernels
if mid:
    cases = [(21, 1), (22, 1), (23, 1), (16, 2), (16, 4), (16, 8), (12, 32), (14, 8)]
for log_n, batch in cases:
    n = batch << log_n
    x = rng.integers(0, 2**31, size=(n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
    d = torch.from_numpy(x).cuda().view(-1)
    rows = []
    for gb in (1, 2, 3):
        for T in range(5, (12 if gb == 3 else 11) + 1):
            if T > log_n: continue
            if mid and not (8 <= T <= 9 if gb == 1 else 9 <= T <= 10 if gb == 2 else T >= 9): continue
            for g in ((0, 2) if mid else (0, 1, 2)):
                ctx = Ntt256("bn254_fr", max_log_n=24)
                try:
                    ctx.set_tuning(T, g); ctx.set_elements_per_thread(gb)
                    us = graph_time(lambda: ctx.forward(d, log_n, batch))
                    rows.append((round(us, 2), gb, T, g))
                except Ntt256Error:
                    pass
                finally:
                    ctx.close()
    rows.sort()
    print(json.dumps(dict(log_n=log_n, batch=batch, best=rows[:6],
                          best_per_gb={gb: min([r for r in rows if r[1] == gb], default=None) for gb in (1, 2, 3)})), flush=True)
