"""384-bit transforms (BLS12-377 Fq) on one GPU: tile sweep at 2^22, then forward / inverse times
over sizes with the default tiling (development aid; profiles/r01_ntt384.jsonl)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt384
from ntt_b200.lib import Ntt256Error


def timeit(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def data(n):
    d = torch.randint(0, 2**31 - 1, (n, 12), dtype=torch.int32, device="cuda")
    d[:, 11] &= 0x00FFFFFF  # < 2^376 < p
    return d.view(-1)


ctx = Ntt384("bls12_377_fq", max_log_n=24)
d = data(1 << 22)
for T in (9, 10, 11):
    for g in (0, 1, 2, 3):
        try:
            ctx.set_tuning(T, g)
            ms = timeit(lambda: ctx.forward(d, 22))
            print(json.dumps(dict(sweep="2^22", T=T, g=g, fwd_ms=round(ms, 4), gelem=round((1 << 22) / ms / 1e6, 3))), flush=True)
        except Ntt256Error as e:
            print(json.dumps(dict(sweep="2^22", T=T, g=g, error=str(e))), flush=True)
del d
ctx.close()
ctx = Ntt384("bls12_377_fq", max_log_n=24)
for log_n, batch in ((12, 1), (16, 1), (20, 1), (22, 1), (24, 1), (16, 256), (12, 4096)):
    n = batch << log_n
    d = data(n)
    f = timeit(lambda: ctx.forward(d, log_n, batch)); i = timeit(lambda: ctx.inverse(d, log_n, batch))
    bf = n / 2 * log_n
    print(json.dumps(dict(log_n=log_n, batch=batch, fwd_ms=round(f, 4), inv_ms=round(i, 4), fwd_gelem=round(n / f / 1e6, 3),
                          g_butterflies_per_s=round(bf / f / 1e6, 2),
                          int_frac_plain_imad=round(bf * 300 / (f / 1e3) / 18.3e12, 3),
                          int_frac_carry_chain=round(bf * 300 / (f / 1e3) / 9.2e12, 3))), flush=True)
    del d
