#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02d_pytest_gpu.log 2>&1; tail -5 $O/r02d_pytest_gpu.log
python scripts/time_shapes.py bn254_fr mid > $O/r02d_shapes_mid_bn254.jsonl 2> $O/r02d_shapes.err; cat $O/r02d_shapes_mid_bn254.jsonl
python scripts/bench_companions.py > $O/r02d_companions.jsonl 2> $O/r02d_companions.err; cat $O/r02d_companions.jsonl; tail -3 $O/r02d_companions.err
