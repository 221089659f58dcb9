#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2q_pytest_gpu.log; tail -4 $OUT/r2q_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/gpu_policy_bench.py 16384 > $OUT/r2q_policy_bench.json; cat $OUT/r2q_policy_bench.json
python tools/gpu_chain_trace.py 2>&1 | grep -v Warn > $OUT/r2q_chain_timeline.txt
timeout 600 python bench.py --workload intention --steps 20 --warmup 3 > $OUT/r2q_bench_intention_1gpu.json 2> $OUT/r2q_bench.err; python -c "
import json; d=json.load(open('$OUT/r2q_bench_intention_1gpu.json')); print('intention', d['value'], d['ms_per_step'], d.get('policy'))"
