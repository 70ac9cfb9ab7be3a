#!/bin/bash
# ncu evidence for the kernels either side of the pass kernel (Pauli sum, fused adjoint step, sampling).
# usage: tools/gpu_round_aux.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
for K in pauli_sum adjoint_step segment_mass sample_resolve; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
      python tools/aux_bench.py 28 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | grep prof_ | tail -6
