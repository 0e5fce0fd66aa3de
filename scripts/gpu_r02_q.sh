#!/bin/bash
# final evidence run of round 2 on one GPU: full GPU suite, both bench arms, ncu launch list of the bench command,
# ncu --set full of the 64-bit pass kernel (after the index tables / shift twiddles)
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02q_pytest_gpu.log 2>&1; tail -4 $O/r02q_pytest_gpu.log
python bench.py > $O/r02q_bench.json 2> $O/r02q_bench.err; cut -c1-600 $O/r02q_bench.json; tail -2 $O/r02q_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02q_bench_ref.json 2> $O/r02q_bench_ref.err; cut -c1-300 $O/r02q_bench_ref.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02q_bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02q_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02q_ncu_bench.log 2>&1
tail -2 $O/r02q_ncu_bench.log | cut -c1-200
python scripts/ncu_target64.py 24 1 2 > $O/r02q_plain64.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt64_pass -s 4 -c 4 -o $O/r02q_pass64_24 python scripts/ncu_target64.py 24 1 2 > $O/r02q_ncu_64.log 2>&1
tail -1 $O/r02q_ncu_64.log
python scripts/ncu_target64.py 12 4096 2 >> $O/r02q_plain64.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt64_pass -s 1 -c 1 -o $O/r02q_pass64_b python scripts/ncu_target64.py 12 4096 2 > $O/r02q_ncu_64b.log 2>&1
tail -1 $O/r02q_ncu_64b.log
