#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_64.py -m gpu -x -q > $O/r02k_pytest64.log 2>&1; tail -5 $O/r02k_pytest64.log
python scripts/bench64.py > $O/r02k_bench64.jsonl 2> $O/r02k_bench64.err; grep goldilocks $O/r02k_bench64.jsonl; tail -3 $O/r02k_bench64.err
