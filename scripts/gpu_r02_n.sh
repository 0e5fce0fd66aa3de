#!/bin/bash
# 8-GPU job: multi-GPU parity suite (2^24 full compare, 2^28 closed form, 384-bit four-step), bench line at N=8, step diagnosis
O=gpurun_out
python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/r02n_pytest_8gpu.log 2>&1; tail -5 $O/r02n_pytest_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02n_bench_8gpu.json 2> $O/r02n_bench_8gpu.err; cat $O/r02n_bench_8gpu.json | cut -c1-3000; tail -3 $O/r02n_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra-configs > $O/r02n_bench_4gpu.json 2> $O/r02n_bench_4gpu.err; cut -c1-400 $O/r02n_bench_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 scripts/diag_fourstep.py > $O/r02n_diag_8gpu.json 2> $O/r02n_diag.err; tail -2 $O/r02n_diag_8gpu.json
