for M in 1 2 3 5; do echo "every $M"; TMJX_SYNC_EVERY=$M timeout 300 python tools/gpu_perf_sweep.py 4096 2>&1 | tail -1; done
