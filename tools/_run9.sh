OUT=gpurun_out; TAG=${1:-r1n}
(
  TMJX_SYNC=2 timeout 200 python tools/gpu_perf_sweep.py 4096
  TMJX_SYNC=1 timeout 200 python tools/gpu_perf_sweep.py 4096
  TMJX_SYNC=0 timeout 200 python tools/gpu_perf_sweep.py 4096
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
