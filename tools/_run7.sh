OUT=gpurun_out; TAG=${1:-r1j}
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
(
  timeout 200 python tools/gpu_perf_sweep.py 2072 4096 8192
  TMJX_ENVS_PER_BLOCK=7 timeout 200 python tools/gpu_perf_sweep.py 2072 4096
  TMJX_NO_GEN=1 timeout 200 python tools/gpu_perf_sweep.py 4096
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
