#!/bin/bash
# One GPU-box session that reproduces the evidence of a round (1 GPU): parity tests, smoke, the bench lines of BASELINE configs[1], [2], [4]
# and [3] (ppo), the residency / phase experiments, the learner kernels, and the ncu launch lists + one full capture of the step kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_ncu]         multi-GPU: tools/gpu_round_multi.sh
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_default_100steps.json 2>> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --full-cpu-baseline > $OUT/${TAG}_bench_full_cpu_baseline.json 2>> $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload intention --steps 20 --warmup 3 > $OUT/${TAG}_bench_intention_1gpu.json 2>> $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload contact --steps 20 --warmup 3 > $OUT/${TAG}_bench_contact_1gpu.json 2>> $OUT/${TAG}_bench.err
timeout 1200 python bench.py --workload ppo --steps 2 --warmup 1 > $OUT/${TAG}_bench_ppo_1gpu.json 2>> $OUT/${TAG}_bench.err
for f in bench bench_default_100steps bench_intention_1gpu bench_contact_1gpu bench_ppo_1gpu; do python -c "
import json,sys; d=json.load(open('$OUT/${TAG}_$f.json')); print('$f', round(d['value']), 'env-steps/s', round(d['ms_per_step'],3), 'ms', (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('episode_stats') or {}))"; done
(
  timeout 300 python tools/gpu_perf_sweep.py 148 2072 4096 8192 16384
  TMJX_L2_SPILL=0 timeout 300 python tools/gpu_perf_sweep.py 4096 16384
  TMJX_ENVS_PER_BLOCK=16 timeout 300 python tools/gpu_perf_sweep.py 2368 4096 16384
  TMJX_NO_GEN=1 timeout 300 python tools/gpu_perf_sweep.py 4096
) > $OUT/${TAG}_sweep.log 2>&1
timeout 600 python tools/gpu_warp_scaling.py > $OUT/${TAG}_warp_scaling.txt 2>&1
if [ -f track-mjx_b200/csrc/libtmjx_pt.so ]; then for k in 1 14; do timeout 300 python tools/gpu_phase_timing.py $k; done > $OUT/${TAG}_phase_timing.txt 2>&1; fi
timeout 300 python tools/gpu_learner_bench.py > $OUT/${TAG}_learner_bench.json 2> $OUT/${TAG}_learner_bench.err
timeout 300 python tools/gpu_policy_bench.py 16384 > $OUT/${TAG}_policy_bench.json 2>> $OUT/${TAG}_learner_bench.err
timeout 300 python tools/gpu_chain_trace.py 2>&1 | grep -v Warn > $OUT/${TAG}_chain_timeline.txt
if [ -z "$2" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches_learner.csv python tools/gpu_learner_bench.py --quick > $OUT/${TAG}_ncu_learner.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_ppo.csv \
      python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 1 --warmup 1 > $OUT/${TAG}_ncu_ppo.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tmjx_env_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT/${TAG}_prof.ncu-rep
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 6 -c 1 -f -o $OUT/${TAG}_chain_prof python tools/gpu_policy_bench.py 16384 > $OUT/${TAG}_chain_ncu.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $OUT/${TAG}_launches_intention.csv \
      python bench.py --workload intention --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch_intention.log 2>&1
fi
