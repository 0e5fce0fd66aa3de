#!/bin/bash
# 64-bit transforms after the index tables: parity suite, timings
O=gpurun_out
python -m pytest tests/test_gpu_64.py -m gpu -x -q > $O/r02j_pytest64.log 2>&1; tail -5 $O/r02j_pytest64.log
python scripts/bench64.py > $O/r02j_bench64.jsonl 2> $O/r02j_bench64.err; cat $O/r02j_bench64.jsonl; tail -3 $O/r02j_bench64.err
