#!/bin/bash
# Round-2 GPU job 3: full GPU test suite, kernel-shape sweep, fused companions.
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02c_pytest_gpu.log 2>&1; tail -5 $O/r02c_pytest_gpu.log
python scripts/time_shapes.py bn254_fr > $O/r02c_shapes_bn254.jsonl 2> $O/r02c_shapes.err; cat $O/r02c_shapes_bn254.jsonl
python scripts/time_shapes.py bls12_381_fr > $O/r02c_shapes_bls.jsonl 2>> $O/r02c_shapes.err; cat $O/r02c_shapes_bls.jsonl
python scripts/bench_companions.py > $O/r02c_companions.jsonl 2> $O/r02c_companions.err; cat $O/r02c_companions.jsonl; tail -3 $O/r02c_companions.err
