#!/bin/bash
# index tables in the wide pass kernels: full GPU suite, timings per twiddle mode, 384-bit and small sizes
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02ad_pytest_gpu.log 2>&1; tail -4 $O/r02ad_pytest_gpu.log
python scripts/time_modes.py > $O/r02ad_time_modes.jsonl 2> $O/r02ad_time_modes.err; cat $O/r02ad_time_modes.jsonl
python scripts/time_384.py 2>/dev/null | grep -v sweep > $O/r02ad_ntt384.jsonl; cat $O/r02ad_ntt384.jsonl | cut -c1-120
