OUT=gpurun_out; TAG=${1:-r1e}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
(
  timeout 300 python tools/gpu_perf_sweep.py 148 1776 3552 4096
  TMJX_ENVS_PER_BLOCK=7 timeout 300 python tools/gpu_perf_sweep.py 2072 4096
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tmjx_env_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_prof.ncu-rep
