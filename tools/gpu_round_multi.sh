#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the scaling lines of BASELINE configs[1] (weak, 4096 envs per GPU) and configs[3] (PPO training,
# 65536 envs split over the GPUs, NCCL gradient all-reduce), launched the way the driver launches bench.py.
# usage: bash tools/gpu_round_multi.sh <tag> <N> [short]      (short: tracking + ppo only)
TAG=${1:-dev}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; echo "tracking rc=$?"
timeout 1200 $RUN --master-port 29512 bench.py --gpus $N --workload ppo --steps 3 --warmup 1 > $OUT/${TAG}_bench_ppo_${N}gpu.json 2> $OUT/${TAG}_bench_ppo_${N}gpu.err; echo "ppo rc=$?"
tail -2 $OUT/${TAG}_bench_ppo_${N}gpu.err
if [ -z "$3" ]; then
timeout 900 $RUN --master-port 29513 bench.py --gpus $N --workload intention --steps 20 --warmup 3 > $OUT/${TAG}_bench_intention_${N}gpu.json 2> $OUT/${TAG}_bench_intention_${N}gpu.err; echo "intention rc=$?"
timeout 900 $RUN --master-port 29514 bench.py --gpus $N --workload contact --steps 20 --warmup 3 > $OUT/${TAG}_bench_contact_${N}gpu.json 2> $OUT/${TAG}_bench_contact_${N}gpu.err; echo "contact rc=$?"
timeout 300 $RUN --master-port 29515 tools/gpu_learner_2gpu.py > $OUT/${TAG}_learner_${N}gpu.json 2> $OUT/${TAG}_learner_${N}gpu.err
fi
for f in bench_${N}gpu bench_ppo_${N}gpu bench_intention_${N}gpu bench_contact_${N}gpu; do [ -s $OUT/${TAG}_$f.json ] && python -c "
import json; d=json.load(open('$OUT/${TAG}_$f.json')); print('$f', round(d['value']), 'env-steps/s', round(d['ms_per_step'],3), 'ms/step', d.get('phases_ms_per_step'), d.get('learner'))"; done
