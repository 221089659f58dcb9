#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2f_pytest.log
tail -30 $OUT/r2f_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5
timeout 900 python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 2 --warmup 1 > $OUT/r2f_bench_ppo_small.json 2> $OUT/r2f_bench_ppo_small.err; echo "bench rc=$?"; tail -3 $OUT/r2f_bench_ppo_small.err
python -c "
import json; d=json.load(open('$OUT/r2f_bench_ppo_small.json')); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['learner'])"
