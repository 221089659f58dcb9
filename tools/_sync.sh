for M in 0 1 2 4 8 16 10 24; do echo "mask $M"; TMJX_SYNC_MASK=$M timeout 300 python tools/gpu_perf_sweep.py 4096 2>&1 | tail -1; done
