#!/bin/bash
# twiddle-array budget test; per-rank four-step arithmetic by kernel shape (8- and 4-GPU shard sizes on one GPU)
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "budget or streamed" > $O/r02g_pytest.log 2>&1; tail -4 $O/r02g_pytest.log
cd scripts && python sweep_fourstep_local.py 24 > ../$O/r02g_fourstep_shapes.jsonl 2> ../$O/r02g_fourstep_shapes.err; cd ..
cat $O/r02g_fourstep_shapes.jsonl; tail -3 $O/r02g_fourstep_shapes.err
