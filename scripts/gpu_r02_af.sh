#!/bin/bash
O=gpurun_out
python scripts/ncu_target.py 24 2 > $O/r02af_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt_pass_kernel -s 4 -c 4 -o $O/r02af_pass python scripts/ncu_target.py 24 2 > $O/r02af_ncu.log 2>&1
tail -2 $O/r02af_ncu.log
