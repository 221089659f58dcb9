#!/bin/bash
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 900 python tools/gpu_warp_scaling.py > $OUT/${TAG}_warp_scaling.txt 2>&1
cat $OUT/${TAG}_warp_scaling.txt
