#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02s_pytest_gpu.log 2>&1; tail -4 $O/r02s_pytest_gpu.log
python scripts/bench64.py > $O/r02s_bench64.jsonl 2> $O/r02s_bench64.err; cat $O/r02s_bench64.jsonl | cut -c1-230; tail -2 $O/r02s_bench64.err
