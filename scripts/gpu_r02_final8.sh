#!/bin/bash
# final multi-GPU bench lines of round 2 (N = 8 / 4 / 2, the defaults)
O=gpurun_out
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$N bench.py --gpus $N --steps 20 --warmup 5 2> $O/r02fin_bench_${N}gpu.err | grep '^{' > $O/r02fin_bench_${N}gpu.json; cut -c1-330 $O/r02fin_bench_${N}gpu.json
done
python -m pytest tests/test_multigpu.py -m gpu -x -q -k "nccl or 64_bit" > $O/r02fin_pytest_8gpu.log 2>&1; tail -2 $O/r02fin_pytest_8gpu.log
