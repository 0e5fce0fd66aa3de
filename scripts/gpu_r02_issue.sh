#!/bin/bash
# Round-2 GPU job 1: instruction-form costs (plain timing), then pipe counters for the same kernels and for one
# ntt_pass_kernel launch.  Run under gpurun from the repo root.
set -x
M="smsp__inst_executed.sum,smsp__inst_issued.sum,smsp__issue_active.sum,smsp__cycles_active.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__pipe_alu_cycles_active.sum,smsp__pipe_fma_cycles_active.sum,smsp__pipe_fmaheavy_cycles_active.sum,smsp__pipe_fmalite_cycles_active.sum,smsp__pipe_fp64_cycles_active.sum,gpu__time_duration.sum"
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $O/r02_mb_clocks.csv &
SMI=$!
./scripts/microbench_issue 4000 5 > $O/r02_mb_issue.jsonl 2> $O/r02_mb_issue.err
kill $SMI
./scripts/microbench_issue 400 1 > $O/r02_mb_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file $O/r02_mb_issue_ncu.csv ./scripts/microbench_issue 400 1 > $O/r02_mb_ncu.log 2>&1
python scripts/ncu_target.py 24 2 > $O/r02_pass_plain.log 2>&1 &&
ncu --metrics $M,breakdown:sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:ntt_pass_kernel -s 1 -c 2 --csv --log-file $O/r02_pass_pipes.csv python scripts/ncu_target.py 24 2 > $O/r02_pass_ncu.log 2>&1
tail -3 $O/r02_mb_issue.jsonl
