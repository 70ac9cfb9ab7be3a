#!/bin/bash
# One GPU visit: parity tests, the bench line, the ncu launch list and one full capture of the pass kernel.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
# launch list of the engine's own kernels (namespace tcb) for one bench step
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:tcb:: -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 8 -c 2 -f -o gpurun_out/prof_pass_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
