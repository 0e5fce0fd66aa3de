"""Ad-hoc GPU correctness + timing sweep (development aid; the judged tests are tests/ -m gpu)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ntt_b200.api import Ntt256, FIELDS
from oracle import c_oracle as co

def check(field, sizes, tunings, max_log_n=20):
    p, g = FIELDS[field]
    ctx = Ntt256(field, max_log_n=max_log_n)
    bad = 0
    for (T, gg) in tunings:
        ctx.set_tuning(T, gg)
        for L in sizes:
            x = co.splitmix_elements(1 << L, p, 1000 + L)
            want = co.ntt(x, p, g, threads=8)
            d = torch.from_numpy(x.view(np.int32)).cuda()
            ctx.forward(d, L)
            torch.cuda.synchronize()
            got = d.cpu().numpy().view(np.uint32)
            okf = np.array_equal(got, want)
            ctx.inverse(d, L)
            torch.cuda.synchronize()
            oki = np.array_equal(d.cpu().numpy().view(np.uint32), x)
            if not (okf and oki):
                bad += 1
                nb = int((got != want).any(axis=1).sum())
                print(f"MISMATCH {field} L={L} T={T} g={gg} fwd={okf} inv={oki} bad_elems={nb}")
    print(f"{field}: {'ALL OK' if not bad else str(bad)+' FAILURES'} sizes={list(sizes)} tunings={tunings}")
    return bad

def bench(field, L, T, gg, max_log_n, reps=5):
    ctx = Ntt256(field, max_log_n=max_log_n)
    ctx.set_tuning(T, gg)
    p, g = FIELDS[field]
    x = co.splitmix_elements(1 << L, p, 7)
    d = torch.from_numpy(x.view(np.int32)).cuda()
    for _ in range(2): ctx.forward(d, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); ctx.forward(d, L); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"bench {field} 2^{L} T={T} g={gg}: {best:.3f} ms  {(1<<L)/best/1e6:.2f} Gelem/s")

if __name__ == "__main__":
    bad = 0
    bad += check("bn254_fr", range(0, 17), [(10, 2), (12, 2), (6, 1), (8, 3)])
    bad += check("bls12_381_fr", [3, 8, 12, 13, 16, 18, 20], [(10, 2), (11, 2), (12, 2), (12, 0)])
    bad += check("ref29", [5, 10, 15], [(10, 2)])
    if not bad:
        for (T, gg) in [(10, 2), (11, 2), (12, 2), (12, 0), (12, 3)]:
            bench("bn254_fr", 24, T, gg, 24)
        bench("bn254_fr", 20, 12, 2, 24)
        bench("bn254_fr", 16, 12, 2, 24)
    sys.exit(1 if bad else 0)
