#!/bin/bash
# round-2 GPU session A: new parity tests on the old and the compact layout, residency sweep, bench line
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
TMJX_L2_SPILL=0 timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_policy.py > $OUT/${TAG}_pytest_oldlayout.log 2>&1; echo "pytest(old layout) rc=$?" | tee -a $OUT/${TAG}_pytest_oldlayout.log
tail -15 $OUT/${TAG}_pytest_oldlayout.log
mkdir -p $OUT/parity_old && mv $OUT/parity_*.json $OUT/parity_old/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest(compact layout) rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
(
  TMJX_L2_SPILL=0 timeout 300 python tools/gpu_perf_sweep.py 148 4096 8192 16384
  TMJX_ENVS_PER_BLOCK=14 timeout 300 python tools/gpu_perf_sweep.py 148 4096 8192 16384
  TMJX_ENVS_PER_BLOCK=16 timeout 300 python tools/gpu_perf_sweep.py 148 2368 4096 8192 16384
  timeout 300 python tools/gpu_perf_sweep.py 4096 16384 32768
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-700 $OUT/${TAG}_bench.json
