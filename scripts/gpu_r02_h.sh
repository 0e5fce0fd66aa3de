#!/bin/bash
# 64-bit transforms: parity suite, timings, tuning sweep
O=gpurun_out
python -m pytest tests/test_gpu_64.py -m gpu -x -q > $O/r02h_pytest64.log 2>&1; tail -15 $O/r02h_pytest64.log
python scripts/bench64.py > $O/r02h_bench64.jsonl 2> $O/r02h_bench64.err; cat $O/r02h_bench64.jsonl; tail -3 $O/r02h_bench64.err
python scripts/bench64.py sweep > $O/r02h_bench64_sweep.jsonl 2>> $O/r02h_bench64.err; grep '"log_n": 24\|"log_n": 16, "batch": 256' $O/r02h_bench64_sweep.jsonl | grep goldilocks
