#!/bin/bash
# Development build of the step library with the per-phase clock64() timers (14-warp CG variant only): track-mjx_b200/csrc/libtmjx_pt.so
set -e
cd "$(dirname "$0")/../track-mjx_b200/csrc"
mkdir -p _obj
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTMJX_PHASE_TIMING"
printf '#define TMJX_VARIANT 14\n#include "../tmjx_step.cu"\n' > _obj/pt_v14.cu
nvcc $F -I . -c -o _obj/pt_v14.o _obj/pt_v14.cu &
nvcc $F -DTMJX_HAVE_VARIANT_14 -c -o _obj/pt_host.o tmjx_step.cu &
nvcc $F -c -o _obj/pt_policy.o tmjx_policy.cu &
nvcc $F -c -o _obj/pt_ffi.o -x cu tmjx_xla_ffi.cc &
wait
nvcc $F -shared -o libtmjx_pt.so _obj/pt_v14.o _obj/pt_host.o _obj/pt_policy.o _obj/pt_ffi.o
ls -la libtmjx_pt.so
