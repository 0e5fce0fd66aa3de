#!/bin/bash
# one pass body for every width: full GPU suite (256 / 384 / 64), timings of the 256- and 384-bit kernels
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02l_pytest_gpu.log 2>&1; tail -6 $O/r02l_pytest_gpu.log
python scripts/time_modes.py > $O/r02l_time_modes.jsonl 2> $O/r02l_time_modes.err; cat $O/r02l_time_modes.jsonl
python scripts/time_384.py > $O/r02l_ntt384.jsonl 2> $O/r02l_ntt384.err; cat $O/r02l_ntt384.jsonl; tail -2 $O/r02l_ntt384.err
