#!/bin/bash
# 8-GPU job after the single-thread system fences: quick parity, bench lines at N = 8 / 4 / 2, step diagnosis
O=gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -x -q -k "nccl or 384" > $O/r02o_pytest_8gpu.log 2>&1; tail -3 $O/r02o_pytest_8gpu.log
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 20 --warmup 5 --no-extra-configs > $O/r02o_bench_${N}gpu.json 2> $O/r02o_bench_${N}gpu.err; cut -c1-330 $O/r02o_bench_${N}gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 scripts/diag_fourstep.py > $O/r02o_diag_8gpu.json 2> $O/r02o_diag.err; tail -1 $O/r02o_diag_8gpu.json
