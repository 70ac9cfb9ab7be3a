#!/bin/bash
# Round-2 evidence visit (one B200): GPU test tier, the driver's bench line, launch lists, ncu captures of the
# pass kernel, the tcgen05 contraction kernel and the adjoint reductions.
TAG=r2
mkdir -p gpurun_out
bash tools/gpu_round.sh $TAG
bash tools/gpu_round_vqe.sh
bash tools/gpu_round_vqe_list.sh
bash tools/gpu_round_tn.sh $TAG
for K in cross_rdm_reg cross_moments; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
      python bench.py --workload vqe --steps 1 --warmup 1 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | grep "_r2\." | tail -20
