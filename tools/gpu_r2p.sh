#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_golden_policy.py tests/test_golden_wrapper.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 3 --warmup 1 > $OUT/r2p_bench_ppo_8192.json 2> $OUT/r2p_bench_ppo_8192.err; echo "rc=$?"; tail -2 $OUT/r2p_bench_ppo_8192.err
python -c "
import json; d=json.load(open('$OUT/r2p_bench_ppo_8192.json')); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['learner']['ms_per_minibatch'], d['losses_last_minibatch'])"
(timeout 200 python tools/gpu_perf_sweep.py 4096 16384; TMJX_LIB_PATH=track-mjx_b200/csrc/libtmjx_inl.so timeout 200 python tools/gpu_perf_sweep.py 4096 16384 2>&1 | sed 's/^/[inline solve] /') 2>&1 | tee $OUT/r2p_inline_solve_sweep.log
