#!/bin/bash
# re-entry check: full GPU suite, bench lines of both arms, timings
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02f_pytest_gpu.log 2>&1; tail -4 $O/r02f_pytest_gpu.log
python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; cat $O/r02f_bench.json; tail -3 $O/r02f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err; cat $O/r02f_bench_ref.json
python scripts/time_modes.py > $O/r02f_time_modes.jsonl 2> $O/r02f_time_modes.err; cat $O/r02f_time_modes.jsonl
