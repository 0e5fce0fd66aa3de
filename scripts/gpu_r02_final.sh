#!/bin/bash
# final evidence run of round 2 on one GPU: full GPU suite, both bench arms, ncu launch list of the bench command,
# ncu --set full of the pass kernel
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02fin_pytest_gpu.log 2>&1; tail -4 $O/r02fin_pytest_gpu.log
python bench.py > $O/r02fin_bench.json 2> $O/r02fin_bench.err; cut -c1-400 $O/r02fin_bench.json; tail -2 $O/r02fin_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02fin_bench_ref.json 2> $O/r02fin_bench_ref.err; cut -c1-200 $O/r02fin_bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02fin_bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02fin_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02fin_ncu_bench.log 2>&1
python scripts/ncu_target.py 24 2 > $O/r02fin_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt_pass -s 4 -c 4 -o $O/r02fin_pass python scripts/ncu_target.py 24 2 > $O/r02fin_ncu.log 2>&1
tail -1 $O/r02fin_ncu.log
python scripts/time_modes.py > $O/r02fin_time_modes.jsonl 2>/dev/null; head -6 $O/r02fin_time_modes.jsonl
