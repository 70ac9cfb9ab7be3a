#!/bin/bash
# generating first pass: parity (whole GPU tier goes through it), A/B of the QAOA evolution, N=1 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py 2>&1 | grep "iter 2\|norm" | sed "s/^/fused /"
TCB_FUSE_START=0 python tools/quick_bench.py 2>&1 | grep "iter 2\|norm" | sed "s/^/init+pass /"
python bench.py --steps 5 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err; tail -2 gpurun_out/bench_gen.err
python - <<PY
import json
for l in open("gpurun_out/bench_gen.json"):
    if l.startswith("{"):
        d = json.loads(l); print("n1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["config"]["cost"], d["e2e"]["cost"], d["gpu_launches"])
PY
