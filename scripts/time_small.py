"""Latency of small transforms: automatic kernel choice (two/four/eight elements per thread) vs the throughput kernel only
(development aid).  Prints one JSON line per size."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256
from ntt_b200.lib import plan_describe


def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 4]


def graph_time(fn, k=20, reps=10):
    """GPU time per call with the host out of the picture: k calls captured in one CUDA graph."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k): fn()
    return timeit(g.replay, reps) / k


a = Ntt256("bn254_fr", max_log_n=24)
rng = np.random.default_rng(1)
for log_n, batch in ((8, 1), (10, 1), (12, 1), (14, 1), (16, 1), (17, 1), (18, 1), (19, 1), (20, 1), (10, 64), (12, 64), (8, 1024)):
    n = batch << log_n
    x = rng.integers(0, 2**31, size=(n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
    d = torch.from_numpy(x).cuda().view(-1)
    row = dict(log_n=log_n, batch=batch)
    for name, limit in (("auto", 1 << 17), ("throughput", 0)):
        a.set_small_transform_limit(limit)
        row[name + "_fwd_us"] = round(timeit(lambda: a.forward(d, log_n, batch)) * 1e3, 2)
        row[name + "_inv_us"] = round(timeit(lambda: a.inverse(d, log_n, batch)) * 1e3, 2)
        row[name + "_fwd_graph_us"] = round(graph_time(lambda: a.forward(d, log_n, batch)) * 1e3, 2)
        row[name + "_inv_graph_us"] = round(graph_time(lambda: a.inverse(d, log_n, batch)) * 1e3, 2)
    print(json.dumps(row), flush=True)
