#!/bin/bash
# quick 1-GPU visit: GPU test tier, the N=1 bench line without sub-records, host-side e2e profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_quick.log
python bench.py --steps 5 --warmup 3 --no-sub-records > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<PY
import json
for l in open("gpurun_out/bench_quick.json"):
    if l.startswith("{"):
        d = json.loads(l); print("n1", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python tools/e2e_prof.py > gpurun_out/e2e_prof_quick.txt 2>&1; head -3 gpurun_out/e2e_prof_quick.txt
