OUT=gpurun_out; TAG=r1d
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
(
  TMJX_ENVS_PER_BLOCK=4 timeout 300 python tools/gpu_perf_sweep.py 148 1776 3552 4096 5328
  TMJX_ENVS_PER_BLOCK=7 timeout 300 python tools/gpu_perf_sweep.py 2072 4096
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
