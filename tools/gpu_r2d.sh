#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > $OUT/r2d_pytest_train.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2d_pytest_train.log
tail -30 $OUT/r2d_pytest_train.log
timeout 900 python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 2 --warmup 1 > $OUT/r2d_bench_ppo_small.json 2> $OUT/r2d_bench_ppo_small.err; echo "bench rc=$?"
tail -5 $OUT/r2d_bench_ppo_small.err; cut -c1-1500 $OUT/r2d_bench_ppo_small.json
