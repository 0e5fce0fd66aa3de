#!/bin/bash
# final multi-GPU evidence of round 2: parity suite on 8 GPUs, bench lines at N = 8 / 4 / 2 with the defaults, step diagnosis
O=gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/r02t_pytest_8gpu.log 2>&1; tail -3 $O/r02t_pytest_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 20 --warmup 5 2> $O/r02t_bench_8gpu.err | grep '^{' > $O/r02t_bench_8gpu.json; cut -c1-330 $O/r02t_bench_8gpu.json
for N in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$N bench.py --gpus $N --steps 20 --warmup 5 2> $O/r02t_bench_${N}gpu.err | grep '^{' > $O/r02t_bench_${N}gpu.json; cut -c1-330 $O/r02t_bench_${N}gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29739 scripts/diag_fourstep.py 2> $O/r02t_diag.err | grep '^{' > $O/r02t_diag_8gpu.json; cat $O/r02t_diag_8gpu.json
