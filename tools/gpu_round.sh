#!/bin/bash
# One GPU-box round: parity tests, variant timing sweep, bench line, ncu launch list + full capture of the step kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_ncu]
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
(
  timeout 300 python tools/gpu_perf_sweep.py 148 2072 4096 8192
  TMJX_NO_GEN=1 timeout 300 python tools/gpu_perf_sweep.py 148 4096
  TMJX_ENVS_PER_BLOCK=4 timeout 300 python tools/gpu_perf_sweep.py 148 1776 4096
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cat $OUT/${TAG}_bench.json | cut -c1-600
timeout 120 python tools/gpu_learner_bench.py > $OUT/${TAG}_learner_bench.json 2> $OUT/${TAG}_learner_bench.err; cut -c1-400 $OUT/${TAG}_learner_bench.json
if [ -z "$2" ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
      --log-file $OUT/${TAG}_launches_learner.csv python tools/gpu_learner_bench.py --quick > $OUT/${TAG}_ncu_learner.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tmjx_env_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT/${TAG}_prof.ncu-rep
fi
