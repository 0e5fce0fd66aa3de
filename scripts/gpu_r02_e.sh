#!/bin/bash
# wide-lazy policy: parity, field microbench, timings, one full ncu capture of the pass kernel
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02e_pytest_gpu.log 2>&1; tail -4 $O/r02e_pytest_gpu.log
./scripts/microbench_field 400 5 > $O/r02e_mb_field.jsonl 2>&1; cat $O/r02e_mb_field.jsonl
python scripts/time_modes.py > $O/r02e_time_modes.jsonl 2> $O/r02e_time_modes.err; cat $O/r02e_time_modes.jsonl
python scripts/ncu_target.py 24 2 > $O/r02e_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt_pass_kernel -s 4 -c 4 -o $O/r02e_pass python scripts/ncu_target.py 24 2 > $O/r02e_ncu.log 2>&1
tail -3 $O/r02e_ncu.log
