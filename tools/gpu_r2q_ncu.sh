#!/bin/bash
# ncu evidence for the fused policy launch: one --set full capture of mlp_chain_kernel + the launch list of the intention workload
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 6 -c 1 -f -o $OUT/r2q_chain_prof python tools/gpu_policy_bench.py 16384 > $OUT/r2q_chain_ncu.log 2>&1
ls -la $OUT/r2q_chain_prof.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r2q_launches_intention.csv python bench.py --workload intention --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2q_ncu_launch.log 2>&1
tail -3 $OUT/r2q_launches_intention.csv | cut -c1-300
