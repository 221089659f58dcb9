OUT=gpurun_out; TAG=r1c
python tools/gpu_debug_flags.py > $OUT/${TAG}_flags.log 2>&1; tail -20 $OUT/${TAG}_flags.log
(
  TMJX_ENVS_PER_BLOCK=7 timeout 300 python tools/gpu_perf_sweep.py 1036 2072
  TMJX_ENVS_PER_BLOCK=4 timeout 300 python tools/gpu_perf_sweep.py 592 1184 1776
  TMJX_ENVS_PER_BLOCK=44 timeout 300 python tools/gpu_perf_sweep.py 592 1184 1776
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
TMJX_ENVS_PER_BLOCK=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tmjx_env_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep
