"""profiles/rNN_ncu_summary.json from an ncu report:
    ncu -i gpurun_out/pass.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_summarize.py /tmp/raw.csv profiles/r01_ncu_summary.json "<how the report was captured>"
Keeps, per launch, the metrics DESIGN.md and bench.py quote (duration, DRAM bytes, pipe utilisation,
occupancy limiters) and the per-launch DRAM traffic in bytes (bench.py's roofline.traffic)."""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__sass_inst_executed_op_local_ld.sum",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(raw_csv, out_json, source):
    rows = [r for r in csv.reader(open(raw_csv)) if r]
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    launches, dram = [], []
    for r in rows[head + 2:]:
        if len(r) != len(names):
            continue
        rec = {"kernel": r[col["Kernel Name"]]}
        for m in KEEP:
            if m in col:
                rec[m] = f"{r[col[m]]} {units[col[m]]}".strip()
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        dram.append(total)
        launches.append(rec)
    out = {
        "source": source,
        "launches": launches,
        "dram_bytes_each_launch": dram,
        "dram_bytes_per_launch": sum(dram) / max(len(dram), 1),
    }
    json.dump(out, open(out_json, "w"), indent=1)
    print(f"{len(launches)} launches, mean DRAM bytes per launch {out['dram_bytes_per_launch']:.4g}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
