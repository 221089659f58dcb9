#!/bin/bash
# Development build of the step library (14-warp CG variant only) with extra -D flags, for A/B runs through TMJX_LIB_PATH:
#   bash tools/build_dev_lib.sh <name> [-DFLAG ...]   ->  track-mjx_b200/csrc/libtmjx_<name>.so
#   bash tools/build_dev_lib.sh pt -DTMJX_PHASE_TIMING          (the per-phase clock64() timers, tools/gpu_phase_timing.py)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../track-mjx_b200/csrc"
mkdir -p _obj
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $*"
printf '#define TMJX_VARIANT 14\n#include "../tmjx_step.cu"\n' > _obj/${NAME}_v14.cu
nvcc $F -I . -c -o _obj/${NAME}_v14.o _obj/${NAME}_v14.cu &
nvcc $F -DTMJX_HAVE_VARIANT_14 -c -o _obj/${NAME}_host.o tmjx_step.cu &
nvcc $F -c -o _obj/${NAME}_policy.o tmjx_policy.cu &
nvcc $F -c -o _obj/${NAME}_ffi.o -x cu tmjx_xla_ffi.cc &
wait
nvcc $F -shared -o libtmjx_${NAME}.so _obj/${NAME}_v14.o _obj/${NAME}_host.o _obj/${NAME}_policy.o _obj/${NAME}_ffi.o
ls -la libtmjx_${NAME}.so
