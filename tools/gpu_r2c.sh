#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for k in 1 14; do timeout 300 python tools/gpu_phase_timing.py $k; done > $OUT/r2c_phase_timing.txt 2>&1
cat $OUT/r2c_phase_timing.txt
