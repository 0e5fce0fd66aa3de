"""ntt256_mgpu_* (one host thread, all GPUs of the box): 2^24 forward, wall time per transform over
20 back-to-back calls bracketed by ntt256_mgpu_synchronize (development aid)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ntt_b200.api import Ntt256MultiGpu

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_dev = torch.cuda.device_count()
for P in (1, 2, 4, 8):
    if P > n_dev: break
    mg = Ntt256MultiGpu(P, "bn254_fr", max_log_n=log_n)
    shards = [torch.randint(0, 2**28, ((1 << log_n) // P * 8,), dtype=torch.int32, device=f"cuda:{i}") for i in range(P)]
    for _ in range(3): mg.forward(shards, log_n)
    mg.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20): mg.forward(shards, log_n)
        mg.synchronize()
        best = min(best, (time.perf_counter() - t0) / 20)
    print(json.dumps(dict(api="ntt256_mgpu_forward", log_n=log_n, gpus=P, ms=round(best * 1e3, 4),
                          g_elements_per_s=round((1 << log_n) / best / 1e9, 2))), flush=True)
    mg.close(); del shards
