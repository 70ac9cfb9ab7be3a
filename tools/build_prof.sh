#!/bin/bash
# Developer build: the library with -DPASS_PROFILE (phase timers / CTA timeline of the pass kernel), used by
# tools/phase_prof.py, tools/pass_trace.py and tools/pass_trace_qaoa.py.
set -e
cd "$(dirname "$0")/../tensorcircuit_ng_b200/csrc"
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC -DPASS_PROFILE \
  -o ../lib/libtcb200_prof.so api.cu pass_kernel.cu sv_kernels.cu tn_kernels.cu tn_gemm_tc.cu plan.cu
echo built ../lib/libtcb200_prof.so
