#!/bin/bash
# programmatic dependent launch: full GPU suite, then sizes with and without it
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_sizes or batches or fourstep or graph or small" > $O/r02r_pytest_gpu.log 2>&1; tail -3 $O/r02r_pytest_gpu.log
python - <<'PY' > $O/r02r_pdl.jsonl 2>&1
import json, sys, os
sys.path.insert(0, os.getcwd())
import torch
from ntt_b200.api import Ntt256
def timed(fn, reps):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def graph_time(fn, k=10, reps=10):
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
    return ts[len(ts) // 4] / k
for field in ("bn254_fr", "bls12_381_fr"):
    ctx = Ntt256(field, max_log_n=24)
    for log_n, batch in ((24, 1), (22, 1), (21, 1), (20, 1), (18, 1), (16, 1), (14, 1), (12, 1), (16, 256), (12, 4096)):
        d = torch.randint(0, 1 << 28, ((batch << log_n) * 8,), dtype=torch.int32, device="cuda")
        row = {"field": field, "log_n": log_n, "batch": batch}
        for pdl in (0, 1):
            ctx.set_dependent_launch(bool(pdl))
            row[f"ms_pdl{pdl}"] = round(timed(lambda: ctx.forward(d, log_n, batch=batch), 30), 5)
            if (batch << log_n) <= (1 << 21):
                try:
                    row[f"graph_ms_pdl{pdl}"] = round(graph_time(lambda: ctx.forward(d, log_n, batch=batch)), 5)
                except Exception as e:
                    row[f"graph_err_pdl{pdl}"] = repr(e)[:120]
        print(json.dumps(row), flush=True)
    ctx.close()
PY
cat $O/r02r_pdl.jsonl
cd scripts && python time_fourstep_local.py 24 2>&1 | grep '"gpus": 8' | head -3
