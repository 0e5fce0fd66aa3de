#!/bin/bash
# ncu --set full of the 64-bit pass kernel: the single-pass batch (pure arithmetic) and the 2^24 transform
O=gpurun_out
python scripts/ncu_target64.py 12 4096 2 > $O/r02i_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt64_pass -s 1 -c 1 -o $O/r02i_pass64_b python scripts/ncu_target64.py 12 4096 2 > $O/r02i_ncu_b.log 2>&1
tail -2 $O/r02i_ncu_b.log
python scripts/ncu_target64.py 24 1 2 >> $O/r02i_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:ntt64_pass -s 4 -c 4 -o $O/r02i_pass64_24 python scripts/ncu_target64.py 24 1 2 > $O/r02i_ncu_24.log 2>&1
tail -2 $O/r02i_ncu_24.log
