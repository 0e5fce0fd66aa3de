"""Per-rank arithmetic of the four-step steps on ONE GPU for every kernel shape (tile size x elements per thread
x the dense four-element instance): which shape fills 148 SMs best when a rank holds 2^(log_n - log_p) elements.
Development aid for the wave-tail question of the 8-GPU strong-scaling run."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt256
from time_fourstep_local import graph_time

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for log_p in (3, 2):
    m = log_n - log_p
    d = torch.randint(0, 2**28, ((1 << m) * 8,), dtype=torch.int32, device="cuda")
    log_r = log_n // 2
    for gb, T, g, dense in ((3, 11, 2, 0), (3, 12, 0, 0), (3, 10, 2, 0), (2, 10, 2, 1), (2, 10, 2, 0), (2, 11, 2, 0), (2, 9, 2, 1)):
        ctx = Ntt256("bn254_fr", max_log_n=log_n)
        ctx.set_tuning(T, g)
        ctx.set_elements_per_thread(gb)
        ctx.set_dense_pass4(bool(dense))
        try:
            s1 = graph_time(lambda: ctx.fourstep_step1(d, log_n, log_r, log_p, 1))
            s2 = graph_time(lambda: ctx.fourstep_step2(d, log_n, log_r, log_p, 1))
            print(json.dumps(dict(log_n=log_n, gpus=1 << log_p, elements_per_thread=1 << gb, T=T, g=g, dense=dense,
                                  step1_us=round(s1, 1), step2_us=round(s2, 1), sum_us=round(s1 + s2, 1))), flush=True)
        except Exception as e:
            print(json.dumps(dict(log_n=log_n, gpus=1 << log_p, gb=gb, T=T, g=g, dense=dense, error=str(e))), flush=True)
        ctx.close()
