"""Timings of the SURVEY par. 8f rows: element-wise product (HBM-bound), Montgomery conversion,
coset transform, and the pipelined host-buffer path for batches (e2e over PCIe)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256, FIELDS

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

ctx = Ntt256("bn254_fr", max_log_n=24)
p, g = FIELDS["bn254_fr"]
rng = np.random.default_rng(1)
n = 1 << 24
def rand(n):
    x = rng.integers(0, 2**31, size=(n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
    return torch.from_numpy(x).cuda().view(-1)
a, b, out = rand(n), rand(n), torch.empty(n * 8, dtype=torch.int32, device="cuda")
peak = 6531.9
for mont in (False, True):
    ms = timeit(lambda: ctx.vec_mul(a, b, out=out, mont_form=mont))
    gbs = n * 96 / ms / 1e6
    print(json.dumps(dict(op="vec_mul", mont_form=mont, n=n, ms=round(ms, 4), gelem_s=round(n / ms / 1e6, 2), hbm_gbs=round(gbs, 1), hbm_frac=round(gbs / peak, 3))))
ms = timeit(lambda: ctx.to_montgomery(a))
print(json.dumps(dict(op="to_montgomery", n=n, ms=round(ms, 4), hbm_gbs=round(n * 64 / ms / 1e6, 1), hbm_frac=round(n * 64 / ms / 1e6 / peak, 3))))
f = timeit(lambda: ctx.forward(a, 24)); fi = timeit(lambda: ctx.inverse(a, 24))
ctx.to_montgomery(b)
for fused in (True, False):
    ctx.set_fused_companions(fused)
    c = timeit(lambda: ctx.coset_forward(a, 24, g)); ci = timeit(lambda: ctx.coset_inverse(a, 24, g))
    fm = timeit(lambda: ctx.forward_mul(a, b, 24)); im = timeit(lambda: ctx.inverse_mul(a, b, 24))
    print(json.dumps(dict(op="companions_2^24", fused=fused, coset_forward_ms=round(c, 4), coset_inverse_ms=round(ci, 4),
                          forward_mul_ms=round(fm, 4), inverse_mul_ms=round(im, 4), plain_forward_ms=round(f, 4),
                          plain_inverse_ms=round(fi, 4), coset_forward_over_plain=round(c / f, 4),
                          forward_mul_over_plain=round(fm / f, 4))))
ctx.set_fused_companions(True)
# small transforms: the fused forms save a launch and an HBM/L2 round trip
for log_n, batch in ((20, 1), (16, 1), (16, 16), (12, 64)):
    m = batch << log_n
    aa, bb = a[: 8 * m], b[: 8 * m]
    row = dict(op="companions_small", log_n=log_n, batch=batch, plain_forward_us=round(1e3 * timeit(lambda: ctx.forward(aa, log_n, batch), 30), 2))
    for fused in (True, False):
        ctx.set_fused_companions(fused)
        row["coset_forward_us_" + ("fused" if fused else "separate")] = round(1e3 * timeit(lambda: ctx.coset_forward(aa, log_n, g, batch=batch), 30), 2)
        row["forward_mul_us_" + ("fused" if fused else "separate")] = round(1e3 * timeit(lambda: ctx.forward_mul(aa, bb, log_n, batch=batch), 30), 2)
    print(json.dumps(row))
ctx.set_fused_companions(True)
del a, b, out
# pipelined host batches: 256 x 2^16 (config 3 of BASELINE.json) through the host entry point
log_n, batch = 16, 256
x = rng.integers(0, 2**31, size=(batch << log_n, 8), dtype=np.int64).astype(np.int32); x[:, 7] &= 0x0FFFFFFF
h_in = torch.from_numpy(x).pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
hi, ho = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)
for _ in range(2): ctx.forward_host(hi, log_n, batch=batch, out=ho)
t0 = time.perf_counter()
for _ in range(5): ctx.forward_host(hi, log_n, batch=batch, out=ho)
dt = (time.perf_counter() - t0) / 5
d = h_in.cuda().view(-1)
k = timeit(lambda: ctx.forward(d, log_n, batch))
print(json.dumps(dict(op="forward_host batch 256 x 2^16 (pinned, chunk-pipelined)", ms=round(dt * 1e3, 3), gelem_s=round((batch << log_n) / dt / 1e9, 3),
                      device_only_ms=round(k, 3), h2d_plus_d2h_bytes=2 * (batch << log_n) * 32,
                      pcie_gbs_each_way=round((batch << log_n) * 32 / dt / 1e9, 1))))
