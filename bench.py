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


def cpu_baseline(sample_log_n: int, threads: int, reps: int = 1):
    """Times the oracle port of the reference's CPU NTT (big-num.cu:37-55 restated over the
    256-bit field) on the host cores.  Returns (elements/s, seconds, description)."""
    import numpy as np  # noqa: F401

    from oracle import c_oracle
    from ntt_b200.api import FIELDS

    p, g = FIELDS[FIELD]
    x = c_oracle.splitmix_elements(1 << sample_log_n, p, 0x4E5454 + 3)
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
    over BN254 Fr -- with every host thread, on a bounded sample (2^22 points per step), and also
    reports the literal reference NTT() (oracle/_ref, int64) for continuity."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    threads = max(1, os.cpu_count() or 1)
    sample_log_n = 22
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
        "config": {"workload": f"forward NTT, {FIELD}, bounded sample 2^{sample_log_n} points per step "
                               f"of the 2^{LOG_N} workload", "field": FIELD},
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


def run_ours(args):
    import numpy as np
    import torch

    from ntt_b200.api import FIELDS, Ntt256
    from oracle import c_oracle  # checker only: the correctness spot-check and the cpu_baseline leg

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

    if world == 1:
        x_host = synthetic_elements(n, p, 0x4E5454 + 3)
        d = torch.from_numpy(x_host.view(np.int32)).cuda()

        def step():
            ctx.forward(d, LOG_N)

        fs = None
    else:
        from ntt_b200.dist import FourStep

        fs = FourStep(ctx, LOG_N, world, rank, fused=not args.no_fused, log_r=(args.log_r or None),
                      barrier=args.barrier)
        x_shard = synthetic_elements(n // world, p, 0x4E5454 + 100 + rank)
        fs.load_shard(torch.from_numpy(x_shard.view(np.int32)).cuda())

        def step():
            fs.forward()

    # ---- correctness spot check before timing (one output via Horner on the CPU oracle)
    if world == 1:
        step()
        torch.cuda.synchronize()
        got = d[: 8 * 4].cpu().numpy().view(np.uint32).reshape(-1, 8)
        for k in ((1, 3) if LOG_N <= 24 else (1,)):
            want = c_oracle.dft_sample(x_host, p, g, k)
            have = sum(int(got[k][i]) << (32 * i) for i in range(8))
            if want != have:
                raise AssertionError(f"bench: X[{k}] differs from the oracle")
        d.copy_(torch.from_numpy(x_host.view(np.int32)))

    for _ in range(args.warmup):
        step()
    clk = ClockSampler(local_rank)  # NVML init takes a different time on every rank: before the barrier
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    barriers0 = fs._epoch if fs is not None else 0  # peer-barrier kernels (fused multi-GPU path)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()  # all ranks enter the timed region together (a late rank would be waited for
        torch.cuda.synchronize()  # inside the first step's barrier and inflate everyone's time)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    clk.poll_until(e1)  # the GPU is still inside the timed region: the host ran ahead of it
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    launches = ctx.launch_count - launches0
    barrier_launches = (fs._epoch if fs is not None else 0) - barriers0
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
        e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32,
               "ms_per_step": dt * 1e3, "api": "ntt256_forward_host (pinned host buffers)"}
    else:
        # every rank: its shard from pinned host memory -> device, distributed transform, result
        # shard back to pinned host memory; max over ranks of the wall time per step
        h_in = torch.from_numpy(x_shard.view(np.int32)).pin_memory().view(-1)
        h_out = torch.empty_like(h_in).pin_memory()
        src = fs.a if not fs.fused else (fs.a if fs._cur == 0 else fs.b)

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

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (ntt_pass_kernel): all launches of a step are that kernel
    hbm_peak, peak_src = _peaks()
    passes = launches / args.steps / (1 if world == 1 else 1)
    launch_ms = ms_total / max(launches, 1)
    n_local = n // world
    alg_bytes_transform = 64.0 * n_local * -(-LOG_N // 12)  # 64 B/element per <=12-stage pass
    alg_bytes_launch = alg_bytes_transform / max(passes, 1)
    achieved = alg_bytes_launch / (launch_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_summary.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    modmuls = (n / 2) * LOG_N
    int_peak_macs = 18.3e12  # measured plain IMAD.WIDE rate, profiles/r01_microbench_pipes.jsonl
    int_achieved = modmuls * MACS_PER_MODMUL / (ms_step / 1e3) / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (256-bit modular integer)", "data": "synthetic",
        "config": {"workload": f"2^{LOG_N}-point forward NTT, {FIELD} (8x u32 limbs), in place, natural order in/out"
                               + ("" if world == 1 else f", four-step over {world} GPUs, exchange = "
                                  + ("NCCL all-to-all" if args.no_fused else "peer stores fused into the butterfly kernel (NVLink), " + args.barrier + " barrier")),
                   "field": FIELD, "log_n": LOG_N, "passes_per_step": passes,
                   "l2_policy": "input 512 MiB > 126 MB L2, no flush needed",
                   "binding_roofline": "int32-multiply pipe (see roofline_int); HBM is ~6x off"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "kernel": "ntt_pass_kernel", "launch_ms": launch_ms,
                     "algorithmic_bytes_per_launch": alg_bytes_launch, "peak_source": peak_src},
        "roofline_int": {"bound": "int32-mul", "achieved": int_achieved / 1e12, "peak": int_peak_macs / 1e12,
                         "unit": "T wide-MAC/s", "frac": int_achieved / int_peak_macs,
                         "work": "136 MACs x (n/2) log2 n modmuls (SURVEY 8d)",
                         "peak_source": "measured plain IMAD.WIDE.U32 rate, profiles/r01_microbench_pipes.jsonl",
                         "carry_chain_peak": 9.2, "frac_of_carry_chain_peak": int_achieved / 9.2e12,
                         "note": "256-bit products need carry-predicate IMAD.WIDE forms, which run at half rate on "
                                 "B200 (9.2 T/s): profiles/r01_microbench_notes.md"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches + barrier_launches),
    }
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
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: NCCL all-to-all instead of fused peer stores")
    ap.add_argument("--barrier", choices=("flags", "nccl"), default="flags",
                    help="multi-GPU fused path: peer-memory flag barrier (default) or a one-element NCCL all-reduce")
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
