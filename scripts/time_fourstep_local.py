"""Local compute of the four-step steps on ONE GPU (no exchange): what a rank of an N-GPU run spends
in kernels, per tile size.  Shows how much of the N-GPU step time is arithmetic (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt256
from ntt_b200.lib import plan_describe_fourstep


def graph_time(fn, k=5, reps=8):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(k): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 4] / k * 1e3


log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for log_p in (1, 2, 3):
    m = log_n - log_p
    d = torch.randint(0, 2**28, ((1 << m) * 8,), dtype=torch.int32, device="cuda")
    log_r = log_n // 2
    for T, g in ((10, 2), (11, 2), (12, 2), (12, 0)):
        ctx = Ntt256("bn254_fr", max_log_n=log_n)
        ctx.set_tuning(T, g)
        try:
            s1 = graph_time(lambda: ctx.fourstep_step1(d, log_n, log_r, log_p, 1))
            s2 = graph_time(lambda: ctx.fourstep_step2(d, log_n, log_r, log_p, 1))
            np1 = len(plan_describe_fourstep(1, log_n, log_r, log_p, 1, log_n, T, g)["passes"])
            np2 = len(plan_describe_fourstep(2, log_n, log_r, log_p, 1, log_n, T, g)["passes"])
            print(json.dumps(dict(log_n=log_n, gpus=1 << log_p, T=T, g=g, step1_us=round(s1, 1), step2_us=round(s2, 1),
                                  sum_us=round(s1 + s2, 1), passes=[np1, np2])), flush=True)
        except Exception as e:
            print(json.dumps(dict(log_n=log_n, gpus=1 << log_p, T=T, g=g, error=str(e))), flush=True)
        ctx.close()
