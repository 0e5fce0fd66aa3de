"""The five BASELINE.json configurations on one GPU (4 and 5: their single-GPU form), device-resident
timing with CUDA events; parity of each is covered by tests/test_gpu_parity.py."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256, FIELDS
from oracle import c_oracle as co
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = []
def dev(field, n, seed):
    return torch.from_numpy(co.splitmix_elements(n, FIELDS[field][0], seed).view(np.int32)).cuda().view(-1)
c1 = Ntt256("bn254_fr", max_log_n=24); d = dev("bn254_fr", 1 << 16, 1)
out.append(dict(config=1, what="2^16 BN254 Fr forward", ms=timeit(lambda: c1.forward(d, 16))))
c2 = Ntt256("bls12_381_fr", max_log_n=20); d = dev("bls12_381_fr", 1 << 20, 2)
out.append(dict(config=2, what="2^20 BLS12-381 Fr forward + inverse", ms=timeit(lambda: (c2.forward(d, 20), c2.inverse(d, 20)))))
d = dev("bn254_fr", 256 << 16, 3)
out.append(dict(config=3, what="256 x 2^16 BN254 Fr forward (one GPU)", ms=timeit(lambda: c1.forward(d, 16, 256))))
d = dev("bn254_fr", 1 << 24, 4)
out.append(dict(config=4, what="2^24 BN254 Fr forward (one GPU)", ms=timeit(lambda: c1.forward(d, 24))))
c1.close(); c5 = Ntt256("bn254_fr", max_log_n=28)
d = torch.randint(0, 2**28, (8 << 28,), dtype=torch.int32, device="cuda")
out.append(dict(config=5, what="2^28 BN254 Fr forward (one GPU, 8 GiB)", ms=timeit(lambda: c5.forward(d, 28), reps=3)))
for o in out:
    n = {1: 1 << 16, 2: 2 << 20, 3: 256 << 16, 4: 1 << 24, 5: 1 << 28}[o["config"]]
    o["g_elements_per_s"] = round(n / o["ms"] / 1e6, 3); o["ms"] = round(o["ms"], 4)
    print(json.dumps(o), flush=True)
