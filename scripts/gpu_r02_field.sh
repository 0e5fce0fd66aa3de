#!/bin/bash
# Round-2 GPU job 2: Montgomery vs Shoup product (known answers + rates) and the clean IMAD.WIDE peak.
set -x
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > $O/r02_field_clocks.csv &
SMI=$!
./scripts/microbench_field 400 5 > $O/r02_mb_field.jsonl 2> $O/r02_mb_field.err
./scripts/microbench_issue 4000 5 k_wide_acc >> $O/r02_mb_field.jsonl
kill $SMI
cat $O/r02_mb_field.jsonl
M="smsp__inst_executed.sum,smsp__issue_active.sum,smsp__cycles_active.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__pipe_alu_cycles_active.sum,smsp__pipe_fmaheavy_cycles_active.sum,gpu__time_duration.sum"
./scripts/microbench_field 100 1 > $O/r02_field_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file $O/r02_mb_field_ncu.csv ./scripts/microbench_field 100 1 > $O/r02_field_ncu.log 2>&1
