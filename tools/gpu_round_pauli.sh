#!/bin/bash
# 1-GPU check of the Pauli-sum kernel: tests, the VQE step, standalone timings (tools/aux_bench.py)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pauli_sum.py tests/test_gpu_parity.py tests/test_torchnn.py -m gpu -x -q > gpurun_out/pytest_pauli.log 2>&1
tail -3 gpurun_out/pytest_pauli.log
timeout 600 python bench.py --workload vqe --steps 3 --warmup 1 > gpurun_out/bench_vqe_r2c.json 2> gpurun_out/bench_vqe_r2c.err
python - <<PY
import json
for l in open("gpurun_out/bench_vqe_r2c.json"):
    if l.startswith("{"):
        d = json.loads(l); print("vqe", d["value"], d["ms_per_step"], d["config"].get("energy_mean"), d["config"].get("grad_norm"), d["roofline"]["frac"])
PY
python tools/aux_bench.py 30 > gpurun_out/aux_bench_r2.txt 2>&1
grep -i "pauli" gpurun_out/aux_bench_r2.txt | head
