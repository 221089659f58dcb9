#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2g_pytest.log
tail -25 $OUT/r2g_pytest.log
timeout 300 python tools/gpu_learner_bench.py > $OUT/r2g_learner_bench.json 2> $OUT/r2g_learner_bench.err; cut -c1-1200 $OUT/r2g_learner_bench.json
TMJX_PPO_TWO_PASS=1 timeout 300 python tools/gpu_learner_bench.py --quick > $OUT/r2g_learner_bench_twopass.json 2>/dev/null; cut -c1-700 $OUT/r2g_learner_bench_twopass.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ppo_ -c 40 --csv --log-file $OUT/r2g_launches_losshead.csv python tools/gpu_learner_bench.py --quick > $OUT/r2g_ncu_learner.log 2>&1
grep -E "ppo_fused|adv_stats" $OUT/r2g_launches_losshead.csv | tail -6
timeout 1200 python bench.py --workload ppo --steps 2 --warmup 1 > $OUT/r2g_bench_ppo_1gpu.json 2> $OUT/r2g_bench_ppo_1gpu.err; echo "ppo bench rc=$?"; tail -3 $OUT/r2g_bench_ppo_1gpu.err
python -c "
import json; d=json.load(open('$OUT/r2g_bench_ppo_1gpu.json')); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['learner'])"
