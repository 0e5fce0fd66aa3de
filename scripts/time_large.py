import json, sys, os
sys.path.insert(0, os.getcwd())
import torch
from ntt_b200.api import Ntt256
def timed(fn, reps):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for log_n in (26, 28):
    ctx = Ntt256("bn254_fr", max_log_n=log_n)
    d = torch.randint(0, 1 << 28, ((1 << log_n) * 8,), dtype=torch.int32, device="cuda")
    row = {"log_n": log_n}
    for name, lim in (("default", None), ("two_lookup", 0), ("streamed_32GiB", 32 << 30)):
        if lim is not None:
            ctx.set_stream_twiddle_limit(lim)
            ctx.set_stream_twiddle_budget(max(lim, 1 << 30))
        row[name + "_ms"] = round(timed(lambda: ctx.forward(d, log_n), 5), 3)
        row[name + "_table_bytes"] = ctx.table_bytes()[1]
    for mode in (False,):
        ctx.set_twiddle_mode(mode)
        ctx.set_stream_twiddle_limit(32 << 30); ctx.set_stream_twiddle_budget(32 << 30)
        row["montgomery_streamed_ms"] = round(timed(lambda: ctx.forward(d, log_n), 5), 3)
    print(json.dumps(row), flush=True)
    ctx.close(); del d; torch.cuda.empty_cache()
