#!/usr/bin/env python
"""bench.py -- headline benchmark of the 256-bit NTT hot path (BASELINE.json metric:
256-bit NTT elements/s at 2^24 on 1/2/4/8 B200, with roofline fractions).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the
                                                             # reference's CPU NTT on host cores
    torchrun ... bench.py --gpus N ...                       # four-step over N GPUs (strong scaling)

A "step" is one forward NTT of 2^24 BN254-Fr elements, in place, natural order in and out.
`value` is elements/s with the input resident in HBM; `e2e` is the same metric through the
host-buffer entry point of the C ABI (ntt256_forward_host: H2D + transform + D2H inside the
timed region).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 24
FIELD = "bn254_fr"
METRIC = "ntt256_elements_per_s_2^24"
UNIT = "elements/s"
MACS_PER_MODMUL = 136  # 2*8^2 + 8 32x32->64 multiply-accumulates (SURVEY.md par. 8d)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


INT_PEAK_NOMINAL = 148 * 4 * 32 / 4 * 1.965e9 / 1e12  # T wide-MAC/s: 4 SMSPs x 32 lanes per SM, 4 clocks per warp instruction


def _int_peak():
    """32x32->64 multiply-accumulate rate of one B200 (T MAC/s).  IMAD.WIDE / IMAD.HI occupy the fmaheavy pipe -- the
    only one with an integer multiplier -- for 4 clocks per warp instruction (ncu pipe counters of
    scripts/microbench_issue.cu, profiles/r02_microbench_issue_ncu.csv), which makes the hardware ceiling
    148 x 4 x 32 / 4 x 1.965 GHz = 9.31 T MAC/s; the best rate a loop of independent multiply-accumulates reached is
    8.63 T/s (profiles/r02_int_peak.json).  The roofline divides by the CEILING: with Shoup products the kernels
    issue 107 multiply-equivalents per modular product where SURVEY 8d's accounting charges 136, so the algorithmic
    rate can exceed what any instruction stream could issue -- dividing by the measured 8.63 would report > 1.
    MEASURED_PEAKS.json has no integer figure."""
    return INT_PEAK_NOMINAL, ("hardware ceiling 148 SMs x 4 SMSPs x 32 lanes / 4 clk x 1.965 GHz (fmaheavy occupancy per "
                              "IMAD.WIDE measured with ncu, profiles/r02_microbench_issue_ncu.csv); best measured "
                              "instruction stream: see peak_measured")


def _int_peak_measured():
    try:
        with open(os.path.join(ROOT, "profiles", "r02_int_peak.json")) as f:
            return float(json.load(f)["wide_mac_T_per_s"])
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons through NVML, sampled while the timed region's kernels run.

    No sampling thread: a Python thread polling NVML competes with the launching thread for the GIL
    and the driver, which costs a multi-GPU step of 0.5 ms (three launches and a barrier) ~25 %.
    Instead the launching thread samples after it has enqueued every timed step, while the GPU is
    still working through them (`poll_until(end_event)`)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def sample(self):
        nv = self._nv
        if nv is None:
            return
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def poll_until(self, end_event, interval_s: float = 0.002):
        """Sample until `end_event` (recorded after the last timed step) has completed."""
        self.sample()  # at least one, taken right after the last launch was enqueued
        while not end_event.query():
            time.sleep(interval_s)
            if not end_event.query():
                self.sample()

    def summary(self):
        s = sorted(self.samples)
        return {
            "sm_mhz": (s[len(s) // 2] if s else None),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


def synthetic_elements(n: int, p: int, seed: int):
    """n pseudo-random canonical field elements as uint32[n, 8] (little-endian limbs): the bench's
    own input generator (numpy), independent of oracle/."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    top = p >> 224  # elements whose top limb is below p's top limb are below p
    x[:, 7] = (x[:, 7].astype(np.uint64) % np.uint64(top)).astype(np.uint32)
    return x


_CPU_INPUT: dict = {}


def cpu_baseline(sample_log_n: int, threads: int, reps: int = 1):
    """Times the oracle port of the reference's CPU NTT (big-num.cu:37-55 restated over the
    256-bit field) on the host cores.  Returns (elements/s, seconds, description)."""
    import numpy as np  # noqa: F401

    from oracle import c_oracle
    from ntt_b200.api import FIELDS

    p, g = FIELDS[FIELD]
    x = _CPU_INPUT.get(sample_log_n)
    if x is None:
        x = _CPU_INPUT[sample_log_n] = c_oracle.splitmix_elements(1 << sample_log_n, p, 0x4E5454 + 3)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        c_oracle.ntt(x, p, g, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return (1 << sample_log_n) / best, best


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference has no 256-bit CPU NTT (its
    NTT() is int64 over a 29-bit prime), so this arm times the oracle port -- the same algorithm
    over BN254 Fr -- with every host thread, on the full 2^24-point workload every step, and also
    reports the literal reference NTT() (oracle/_ref, int64) for continuity."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    threads = max(1, os.cpu_count() or 1)
    sample_log_n = LOG_N  # the full workload: ~2 s per step on 16 cores, so K steps still end within minutes
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(sample_log_n, threads)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_baseline(sample_log_n, threads)
        t_total += dt
    ms = t_total / args.steps * 1e3
    value = (1 << sample_log_n) / (ms / 1e3)
    extra = {}
    ref = c_oracle.ref_lib()
    if ref is not None:
        import ctypes

        import numpy as np

        a = np.arange(1 << 20, dtype=np.int64)
        t0 = time.perf_counter()
        ref.ref_ntt_i64(a.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 20)
        extra["reference_i64_ntt"] = {
            "value": (1 << 20) / (time.perf_counter() - t0), "unit": UNIT, "cores": 1,
            "what": "reference NTT() compiled from /root/reference (int64, p=469762049), 2^20 points",
        }
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (256-bit modular integer)", "data": "synthetic",
        "config": {"workload": f"2^{LOG_N}-point forward NTT, {FIELD} (8x u32 limbs), in place, natural order in/out",
                   "field": FIELD, "log_n": LOG_N, "sample": "the whole workload every step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"2^{sample_log_n}-point forward NTT per step, OpenMP over butterflies"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # the reference's own GPU kernels on this box, for the same-hardware comparison (extra key; the
    # arm's value stays the CPU path, as the contract asks)
    try:
        import torch

        rows = _reference_gpu_rows(300) if torch.cuda.is_available() else None
    except Exception:
        rows = None
    if rows:
        r24 = rows[-1]
        extra["reference_gpu_kernels_same_box"] = {
            "what": "oracle/_ref/ref_bignum = /root/reference/src/big-num.cu compiled unmodified for sm_100a; kernel-only "
                    "cudaEvent ms as the program prints them, 256-bit containers, x[i] = i",
            "log_n": r24["log_n"], "naive_ms": r24["naive"], "gzkp_ms": r24["GZKP"],
            "elements_per_s": (1 << r24["log_n"]) / (min(r24["naive"], r24["GZKP"]) / 1e3),
        }
    line.update(extra)
    print(json.dumps(line), flush=True)


def _reference_gpu_rows(timeout_s: int = 600):
    """Runs oracle/_ref/ref_bignum (the reference's 256-bit GPU program, unmodified) and parses its
    printout into one dict per size (2^5..2^24): {log_n, cpu, naive, GZKP} in ms; None if unavailable."""
    import re
    import subprocess

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bignum")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=timeout_s).stdout
    except Exception:
        return None
    rows, cur = [], {}
    for ln in out.splitlines():
        m = re.match(r"(cpu|naive|GZKP): ([0-9.]+)ms", ln.strip())
        if not m:
            cur["mismatch_lines"] = cur.get("mismatch_lines", 0) + 1
            continue
        cur[m.group(1)] = float(m.group(2))
        if m.group(1) == "GZKP":
            rows.append(cur)
            cur = {}
    for i, r in enumerate(rows):
        r["log_n"] = 5 + i  # the program sweeps kk = 5..24 (big-num.cu:370)
    if not rows or min(rows[-1].get("naive", 0.0), rows[-1].get("GZKP", 0.0)) <= 0.0:
        return None
    return rows


def run_reference_gpu(args):
    """--impl reference-gpu: the reference's OWN GPU implementation on this box.  oracle/_ref/ref_bignum
    is /root/reference/src/big-num.cu compiled unmodified for sm_100a (oracle/Makefile, target ref256):
    its main() runs NTT_GPU_Naive and NTT_GZKP (CGBN, 8 threads per 256-bit element) on x[i] = i for
    2^5..2^24 points, checks each against its CPU NTT and prints kernel-only cudaEvent times -- the
    same timing convention as this bench's `value`.  The 256-bit arithmetic does not depend on the
    modulus, so its 2^24 line is comparable with the BN254 workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    rows = _reference_gpu_rows()
    if rows is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/ref_bignum is not built (make -C oracle ref256 needs "
                          "/root/reference) or produced no kernel timings (no CUDA device?)"}))
        return
    last = rows[-1]
    best_ms = min(last["naive"], last["GZKP"])
    n = 1 << last["log_n"]
    line = {
        "impl": "reference-gpu", "metric": f"ntt256_elements_per_s_2^{last['log_n']}", "value": n / (best_ms / 1e3),
        "unit": UNIT, "n_gpus": 1, "ms_per_step": best_ms, "higher_is_better": True, "dtype": "u32x8 (CGBN, 8 threads per element)",
        "data": "x[i] = i (the program's own input)",
        "config": {"workload": f"2^{last['log_n']}-point forward NTT, 256-bit containers, p = 469762049, unmodified "
                               "reference program (NTT_GPU_Naive / NTT_GZKP, kernel-only cudaEvent time as it prints it)"},
        "reference_ms": {"naive": last["naive"], "gzkp": last["GZKP"], "cpu_int64_single_thread": last["cpu"]},
        "self_check_mismatch_lines": sum(r.get("mismatch_lines", 0) for r in rows),
        "all_sizes": [{k: r[k] for k in ("log_n", "cpu", "naive", "GZKP")} for r in rows],
    }
    print(json.dumps(line), flush=True)


def _device_elements(torch, n: int, seed: int):
    """n canonical elements made on the device (int32 limbs, top limb < 2^28 < p's top limb)."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    x = torch.randint(0, 1 << 31, (n, 8), dtype=torch.int32, device="cuda", generator=gen)
    x[:, 7] &= 0x0FFFFFFF
    return x.view(-1)


def _timed(torch, dist, stream, step, steps, clk=None):
    """Device time of `steps` calls of step(): barrier + synchronize on both sides, CUDA events on the
    launching stream, max over ranks.  Returns ms in total."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    if clk is not None:
        clk.poll_until(e1)  # the GPU is still inside the timed region: the host ran ahead of it
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def _mgpu_parity(torch, dist, np, ctx, args, world, rank, p, g):
    """Before anything is timed on N > 1 GPUs: a 2^20-point four-step transform through the same code
    path (same exchange, same barrier), gathered on rank 0 and compared element by element with the
    oracle (/root/reference/src/big-num.cu:446-467 checks every size it runs).  Returns a description."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle

    log_n = 20
    shards = [synthetic_elements((1 << log_n) // world, p, 7000 + r) for r in range(world)] if rank == 0 else None
    mine = synthetic_elements((1 << log_n) // world, p, 7000 + rank)
    fs = nd.FourStep(ctx, log_n, world, rank, fused=not args.no_fused, barrier=args.barrier)
    fs.load_shard(torch.from_numpy(mine.view(np.int32)).cuda().view(-1))
    out = fs.forward().clone()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    ok = True
    if rank == 0:
        X = nd.gather_output([t.cpu().numpy().view(np.uint32).reshape(-1, 8) for t in gathered], log_n, world)
        x = nd.gather_input_layout(shards, log_n, world)
        ok = bool(np.array_equal(X, c_oracle.ntt(x, p, g, threads=max(1, os.cpu_count() or 1))))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    fs.close()
    if not int(flag.item()):
        raise AssertionError(f"bench: four-step 2^{log_n} over {world} GPUs differs from the oracle")
    return f"2^{log_n} four-step over {world} ranks == oracle (all {1 << log_n} outputs)"


def _mgpu_spot_check(torch, dist, np, fs, x_shard, world, rank, p, g, ks):
    """Horner check of outputs X[k], k in ks, of the LOG_N-point distributed transform that fs has just
    run on the shards made by synthetic_elements(seed 100 + rank)."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle

    n = 1 << LOG_N
    out = fs.a if not fs.fused else (fs.a if fs._cur == 0 else fs.b)  # the buffer holding the result
    vals = []
    for k in ks:
        owner, idx = nd.output_location(k, LOG_N, world, fs.log_r)
        t = torch.zeros(8, dtype=torch.int32, device="cuda")
        if rank == owner:
            t.copy_(out.view(-1)[8 * idx: 8 * idx + 8])
        dist.broadcast(t, owner)
        vals.append(t.cpu().numpy().view(np.uint32))
    if rank == 0:
        x = nd.gather_input_layout([synthetic_elements(n // world, p, 0x4E5454 + 100 + r) for r in range(world)],
                                   LOG_N, world, fs.log_r)
        for k, v in zip(ks, vals):
            have = sum(int(v[i]) << (32 * i) for i in range(8))
            if c_oracle.dft_sample(x, p, g, k) != have:
                raise AssertionError(f"bench: X[{k}] of the 2^{LOG_N} transform over {world} GPUs differs from the oracle")
    dist.barrier()
    return f"X[k] of the timed 2^{LOG_N} transform == Horner on the CPU for k in {list(ks)}"


def _config3(torch, dist, np, ctx, args, world, rank, p, g, stream):
    """BASELINE config 3: a batch of 256 independent 2^16-point transforms, sharded over the ranks with no
    collective (ntt_b200.dist.shard_batch); transform 0 of every rank is compared with the oracle."""
    from ntt_b200 import dist as nd
    from oracle import c_oracle

    batch, log_n = 256, 16
    lo, hi = nd.shard_batch(batch, world, rank)
    d = _device_elements(torch, (hi - lo) << log_n, 3000 + rank)
    first = d[: 8 << log_n].cpu().numpy().view(np.uint32).reshape(-1, 8).copy()
    ctx.forward(d, log_n, batch=hi - lo)
    torch.cuda.synchronize()
    got = d[: 8 << log_n].cpu().numpy().view(np.uint32).reshape(-1, 8)
    if not np.array_equal(got, c_oracle.ntt(first, p, g, threads=4)):
        raise AssertionError("bench: config 3 transform differs from the oracle")
    for _ in range(2):
        ctx.forward(d, log_n, batch=hi - lo)
    steps = max(5, min(args.steps, 20))
    ms = _timed(torch, dist, stream, lambda: ctx.forward(d, log_n, batch=hi - lo), steps) / steps
    return {"workload": "256 x 2^16-point forward NTTs, bn254_fr, sharded %d per GPU, no collective" % (hi - lo),
            "ms_per_step": ms, "value": (batch << log_n) / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "checked": "transform 0 of every rank == oracle"}


def _config5(torch, dist, np, args, world, rank, local_rank, p, g, stream):
    """BASELINE config 5: one 2^28-point transform (8 GiB) four-step over 8 GPUs.  Input x[j] = c*h^j (built on
    the device from two small power tables), so sampled outputs are checked against the closed form."""
    from ntt_b200 import dist as nd
    from ntt_b200.api import Ntt256

    log_n = 28
    n = 1 << log_n
    c, h = 0x1234567890ABCDEF1122334455667788 % p, 7
    ctx = Ntt256(FIELD, max_log_n=log_n, device=local_rank)
    fs = nd.FourStep(ctx, log_n, world, rank, fused=not args.no_fused, barrier=args.barrier)
    col, row = nd.geometric_shard(c, h, p, log_n, world, rank, fs.log_r)
    R, w = col.shape[0], row.shape[0]
    a = torch.from_numpy(col.view(np.int32)).cuda().view(R, 1, 8).expand(R, w, 8).contiguous().view(-1)
    b = torch.from_numpy(row.view(np.int32)).cuda().view(1, w, 8).expand(R, w, 8).contiguous().view(-1)
    ctx.vec_mul(a, b, out=a)  # shard[j1][j2] = col[j1] * row[j2]
    del b
    fs.load_shard(a)
    del a
    torch.cuda.empty_cache()
    fs.forward()
    torch.cuda.synchronize()
    out = fs.a if fs._cur == 0 else fs.b
    bad = 0
    for k in (0, 1, n // 2 + 7, n - 1, 123456789 % n, (rank * 0x9E3779B1 + 17) % n):
        owner, idx = nd.output_location(k, log_n, world, fs.log_r)
        if owner == rank:
            v = out.view(-1)[8 * idx: 8 * idx + 8].cpu().numpy().view(np.uint32)
            bad += int(sum(int(v[i]) << (32 * i) for i in range(8)) != nd.geometric_output(c, h, p, g, log_n, k))
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if int(t.item()):
        raise AssertionError("bench: 2^28 transform over 8 GPUs differs from the closed form")
    fs.forward()
    steps = 3
    ms = _timed(torch, dist, stream, fs.forward, steps) / steps
    nvlink_bytes = 32.0 * (n / world) * (world - 1) / world  # sent (= received) per GPU per transform
    # the same transform on ONE B200 (it fits: 8 GiB of data), measured by scripts/time_large.py on this pool
    one_gpu_ms, one_gpu_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_large_1gpu.jsonl")) as f:
            for line in f:
                row = json.loads(line)
                if row.get("log_n") == log_n:
                    one_gpu_ms = min(v for k, v in row.items() if k.endswith("_ms") and "montgomery" not in k)
                    one_gpu_src = "profiles/r02_large_1gpu.jsonl (scripts/time_large.py, one B200; not re-measured by this run)"
    except Exception:
        pass
    fs.close()
    ctx.close()
    return {"workload": "2^28-point forward NTT, bn254_fr, four-step over 8 GPUs (1 GiB shard per GPU)",
            "ms_per_step": ms, "value": n / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "nvlink_bytes_per_gpu_per_step": nvlink_bytes,
            "nvlink_floor_ms": nvlink_bytes / 770e9 * 1e3,
            "int_floor_ms": (n / 2) * log_n * MACS_PER_MODMUL / world / (_int_peak()[0] * 1e12) * 1e3,
            "single_gpu_ms": one_gpu_ms, "single_gpu_source": one_gpu_src,
            "efficiency_vs_single_gpu": (one_gpu_ms / world / ms) if one_gpu_ms else None,
            "checked": "sampled X[k] == c(h^n-1)/(h w^k-1) on every rank"}


def _other_widths(torch, local_rank, steps):
    """SURVEY 8f row 4 made driver-visible (N = 1 only): the 64-bit (Goldilocks) and 384-bit (BLS12-377 Fq)
    instances of the same path, device-timed on synthetic canonical data, each with the roofline figures that
    bound it.  Correctness of these widths is the GPU test suite's job (tests/test_gpu_64.py, test_gpu_384.py);
    here X[0] = sum x is checked for the 64-bit transform as a cheap guard."""
    from ntt_b200.api import Ntt384
    from ntt_b200.api64 import Ntt64, plan_describe64

    hbm_peak, _ = _peaks()
    int_peak, _ = _int_peak()
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    log_n = 24
    p64 = 0xFFFFFFFF00000001
    c64 = Ntt64("goldilocks", max_log_n=log_n, device=local_rank)
    d = torch.randint(0, 1 << 62, (1 << log_n,), dtype=torch.int64, device="cuda")
    lo = int((d & 0xFFFFFFFF).sum().item())
    hi = int((d >> 32).sum().item())
    want0 = (lo + (hi << 32)) % p64
    chk = d.clone()
    c64.forward(chk, log_n)
    got0 = int(chk[0].item()) & ((1 << 64) - 1)
    if got0 != want0:
        raise AssertionError("bench: 64-bit X[0] != sum of the inputs")
    del chk
    launches0 = c64.launch_count
    ms = timed(lambda: c64.forward(d, log_n))
    passes = (c64.launch_count - launches0) / (steps + 3)
    n = 1 << log_n
    out["goldilocks_64bit"] = {
        "workload": f"2^{log_n}-point forward NTT, p = 2^64 - 2^32 + 1, one u64 per element", "ms_per_step": ms,
        "value": n / ms * 1e3, "unit": UNIT, "G_butterflies_s": n / 2 * log_n / ms / 1e6, "passes_per_step": passes,
        "roofline_hbm": {"achieved": 16.0 * n * passes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                         "frac": 16.0 * n * passes / ms / 1e6 / hbm_peak, "bytes": "16 B per element per pass"},
        "bound": "integer ALU issue (add/sub/reduce of 64-bit residues on 32-bit lanes: ~36 instructions per "
                 "butterfly, ALU pipe 64-72 % busy, HBM 17-47 % per pass: profiles/r02_ncu_summary_64.json)",
        "checked": "X[0] == sum x mod p"}
    c64.close()
    del d
    log_n = 22
    c384 = Ntt384("bls12_377_fq", max_log_n=log_n, device=local_rank)
    d = torch.randint(0, 2**31 - 1, (1 << log_n, 12), dtype=torch.int32, device="cuda")
    d[:, 11] &= 0x00FFFFFF
    d = d.view(-1)
    ms = timed(lambda: c384.forward(d, log_n))
    n = 1 << log_n
    macs = (n / 2) * log_n * (2 * 12 * 12 + 12)
    out["bls12_377_fq_384bit"] = {
        "workload": f"2^{log_n}-point forward NTT, 12 x u32 limbs per element", "ms_per_step": ms, "value": n / ms * 1e3,
        "unit": UNIT, "roofline": {"bound": "int32-mul", "achieved": macs / ms / 1e9, "peak": int_peak, "unit": "T wide-MAC/s",
                                   "frac": macs / ms / 1e9 / int_peak, "work": "300 MACs x (n/2) log2 n modmuls"}}
    c384.close()
    return out


def run_ours(args):
    import numpy as np
    import torch

    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle  # checker only: the correctness checks and the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the NTT has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    p, g = FIELDS[FIELD]
    n = 1 << LOG_N
    ctx = Ntt256(FIELD, max_log_n=LOG_N, device=local_rank)
    if args.tile_bits:
        ctx.set_tuning(args.tile_bits, args.coalesce_bits)
    stream = torch.cuda.current_stream()
    checks = []

    if world == 1:
        x_host = synthetic_elements(n, p, 0x4E5454 + 3)
        d = torch.from_numpy(x_host.view(np.int32)).cuda()

        def step():
            ctx.forward(d, LOG_N)

        fs = None
    else:
        from ntt_b200.dist import FourStep

        checks.append(_mgpu_parity(torch, dist, np, ctx, args, world, rank, p, g))
        fs = FourStep(ctx, LOG_N, world, rank, fused=not args.no_fused, log_r=(args.log_r or None),
                      barrier=args.barrier)
        x_shard = synthetic_elements(n // world, p, 0x4E5454 + 100 + rank)
        fs.load_shard(torch.from_numpy(x_shard.view(np.int32)).cuda())

        def step():
            fs.forward()

    # ---- correctness before timing: outputs of the workload itself by Horner on the CPU oracle
    step()
    torch.cuda.synchronize()
    if world == 1:
        got = d[: 8 * 4].cpu().numpy().view(np.uint32).reshape(-1, 8)
        for k in ((1, 3) if LOG_N <= 24 else (1,)):
            want = c_oracle.dft_sample(x_host, p, g, k)
            have = sum(int(got[k][i]) << (32 * i) for i in range(8))
            if want != have:
                raise AssertionError(f"bench: X[{k}] differs from the oracle")
        checks.append(f"X[1], X[3] of the timed 2^{LOG_N} transform == Horner on the CPU")
        d.copy_(torch.from_numpy(x_host.view(np.int32)))
    elif LOG_N <= 24:
        checks.append(_mgpu_spot_check(torch, dist, np, fs, x_shard, world, rank, p, g, (1, n // 2 + 7)))

    for _ in range(args.warmup):
        step()
    clk = ClockSampler(local_rank)  # NVML init takes a different time on every rank: before the barrier
    launches0 = ctx.launch_count
    barriers0 = fs._epoch if fs is not None else 0  # peer-barrier kernels (fused multi-GPU path)
    ms_total = _timed(torch, dist, stream, step, args.steps, clk)
    if fs is not None:
        fs.check()  # no rank-synchronisation wait timed out
    launches = ctx.launch_count - launches0
    barrier_launches = ((fs._epoch if fs is not None else 0) - barriers0) if args.barrier == "flags" else 0
    ms_step = ms_total / args.steps
    value = n / (ms_step / 1e3)

    # ---- end to end through the host-buffer entry point (H2D + NTT + D2H in the timed region)
    e2e = None
    if world == 1:
        h_in = torch.from_numpy(x_host.view(np.int32)).pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        hin_np, hout_np = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)
        for _ in range(2):
            ctx.forward_host(hin_np, LOG_N, out=hout_np)
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.forward_host(hin_np, LOG_N, out=hout_np)
        dt = (time.perf_counter() - t0) / e2e_steps
        # the floor of any host-buffer call: the two PCIe copies alone (a single transform cannot overlap
        # them: every output depends on every input)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            d.copy_(h_in, non_blocking=True)
            h_out.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
        floor = (time.perf_counter() - t0) / 2
        e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
               "ms_per_step": dt * 1e3, "floor_ms": floor * 1e3,
               "floor": "H2D + D2H of the same pinned buffers with no transform (PCIe-bound)",
               "api": "ntt256_forward_host (pinned host buffers, zero-copy first/last pass)"}
    else:
        # every rank: its shard from pinned host memory -> device, distributed transform, result
        # shard back to pinned host memory; max over ranks of the wall time per step
        h_in = torch.from_numpy(x_shard.view(np.int32)).pin_memory().view(-1)
        h_out = torch.empty_like(h_in).pin_memory()

        def e2e_step():
            cur = fs.a if not fs.fused else (fs.a if fs._cur == 0 else fs.b)
            cur.copy_(h_in, non_blocking=True)
            out = fs.forward()
            h_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        dist.barrier()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
               "ms_per_step": dt * 1e3, "api": "per-rank pinned shard -> FourStep.forward -> pinned shard (all ranks)"}

    # ---- the other multi-GPU deliverables of BASELINE.json, as extra keys (every rank takes part)
    extra = {}
    if not args.no_extra_configs and LOG_N == 24:
        try:
            extra["config3"] = _config3(torch, dist, np, ctx, args, world, rank, p, g, stream)
        except AssertionError:
            raise
        except Exception as e:  # an extra key must never cost the headline line
            extra["config3"] = {"error": repr(e)[:200]}
        if world == 1:
            try:
                extra["other_widths"] = _other_widths(torch, local_rank, max(3, min(args.steps, 10)))
            except AssertionError:
                raise
            except Exception as e:
                extra["other_widths"] = {"error": repr(e)[:200]}
        if world == 8 and FIELD == "bn254_fr":
            try:
                if fs is not None:
                    fs.close()
                    fs = None
                extra["config5"] = _config5(torch, dist, np, args, world, rank, local_rank, p, g, stream)
            except AssertionError:
                raise
            except Exception as e:
                extra["config5"] = {"error": repr(e)[:200]}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- rooflines of the dominant kernel (ntt_pass_kernel): all launches of a step are that kernel
    hbm_peak, hbm_src = _peaks()
    int_peak, int_src = _int_peak()
    passes = launches / args.steps
    launch_ms = ms_total / max(launches, 1)
    n_local = n // world
    alg_bytes_transform = 64.0 * n_local * -(-LOG_N // 12)  # 64 B/element per <=12-stage pass
    alg_bytes_launch = alg_bytes_transform / max(passes, 1)
    hbm_achieved = alg_bytes_launch / (launch_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    if world == 1 and LOG_N == 24:
        for name in ("r02_ncu_summary_final.json", "r02_ncu_summary.json", "r01_ncu_summary.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                try:
                    with open(tpath) as f:
                        traffic = json.load(f).get("dram_bytes_per_launch")
                    traffic_src = f"profiles/{name}: ncu --set full of this workload at N=1 (dram__bytes_read+write per launch); not re-measured by this run"
                    break
                except Exception:
                    traffic = None
    modmuls = (n / 2) * LOG_N
    macs_launch = modmuls * MACS_PER_MODMUL / world / max(passes, 1)
    int_achieved = macs_launch / (launch_ms / 1e3) / 1e12
    issued = 107.0 if ctx.twiddle_mode == "shoup" else 132.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (256-bit modular integer)", "data": "synthetic",
        "config": {"workload": f"2^{LOG_N}-point forward NTT, {FIELD} (8x u32 limbs), in place, natural order in/out"
                               + ("" if world == 1 else f", four-step over {world} GPUs, exchange = "
                                  + ("NCCL all-to-all" if args.no_fused else "peer stores fused into the butterfly kernel (NVLink), " + args.barrier + " barrier")),
                   "field": FIELD, "log_n": LOG_N, "passes_per_step": passes,
                   "l2_policy": "input 512 MiB > 126 MB L2, no flush needed",
                   "checked_before_timing": checks},
        # the binding roofline: the 32x32->64 integer multiplier (fmaheavy pipe); HBM is ~6x off
        "roofline": {"bound": "int32-mul", "achieved": int_achieved, "peak": int_peak, "unit": "T wide-MAC/s",
                     "frac": int_achieved / int_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "ntt_pass_kernel", "launch_ms": launch_ms,
                     "algorithmic_macs_per_launch": macs_launch,
                     "work": "136 MACs x (n/2) log2 n modmuls (SURVEY 8d), whatever the multiplier really issues",
                     # what the kernel really issues per product, in units of one wide multiply (4 fmaheavy clocks;
                     # a low 32-bit multiply takes 2): Shoup product 99 + 16/2, Montgomery product 128 + 8/2
                     "peak_measured": _int_peak_measured(),
                     "frac_of_peak_measured": (int_achieved / _int_peak_measured()) if _int_peak_measured() else None,
                     "peak_measured_source": "profiles/r02_int_peak.json: scripts/microbench_issue.cu k_wide_acc, best "
                                             "sustained rate of independent IMAD.WIDE multiply-accumulates on one B200",
                     "twiddle_mode": ctx.twiddle_mode,
                     "issued_wide_equivalents_per_modmul": issued,
                     "frac_issued": int_achieved / int_peak * issued / MACS_PER_MODMUL,
                     "frac_issued_note": "share of the multiplier ceiling taken by the multiplies the kernel "
                                         "executes (ncu: fmaheavy pipe 70 % busy, profiles/r02_ncu_stalls_by_opcode.md); "
                                         "frac above is SURVEY 8d's algorithmic accounting",
                     "peak_source": int_src,
                     "frac_of_r01_denominator": int_achieved / 18.3,
                     "r01_denominator_note": "round 1 divided by 18.3 T/s, a figure its microbenchmark never measured "
                                             "(ptxas had hoisted the product: the loop timed 64-bit adds); see BASELINE.md"},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_achieved / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "peak_source": hbm_src},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches + barrier_launches),
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        threads = max(1, os.cpu_count() or 1)
        v, secs = cpu_baseline(23, threads)
        v1, secs1 = cpu_baseline(19, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"2^23-point forward NTT of the same field on {threads} threads, {secs:.2f} s wall; "
                                          f"single thread (the reference's convention) 2^19 points: {v1:.0f} elements/s"}
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    global LOG_N, METRIC, FIELD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--coalesce-bits", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config3 / config5 extra keys")
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: NCCL all-to-all instead of fused peer stores")
    ap.add_argument("--barrier", choices=("kernel", "flags", "nccl"), default="flags",
                    help="multi-GPU fused path: a one-warp peer-memory barrier kernel between the steps (default; fastest at "
                         "8 GPUs, profiles/r02_diag_8gpu_b.json), synchronisation folded into the pass kernels, or a "
                         "one-element NCCL all-reduce")
    ap.add_argument("--log-r", type=int, default=0, help="multi-GPU: rows = 2^log_r of the four-step view (0: automatic)")
    ap.add_argument("--field", default="bn254_fr", choices=["bn254_fr", "bls12_381_fr"])
    ap.add_argument("--log-n", type=int, default=24, help="transform size (default: the BASELINE metric's 2^24)")
    args = ap.parse_args()
    FIELD = args.field
    if args.log_n != 24:
        LOG_N = args.log_n
        METRIC = f"ntt256_elements_per_s_2^{LOG_N}"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
