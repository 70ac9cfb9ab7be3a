#!/bin/bash
# 1-GPU check of the adjoint reductions: kernel tests, gradient parity, the VQE value_and_grad step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_plans.py tests/test_torchnn.py -m gpu -x -q > gpurun_out/pytest_vqe.log 2>&1
tail -3 gpurun_out/pytest_vqe.log
timeout 600 python bench.py --workload vqe --steps 3 --warmup 1 > gpurun_out/bench_vqe_r2b.json 2> gpurun_out/bench_vqe_r2b.err
python - <<PY
import json
for l in open("gpurun_out/bench_vqe_r2b.json"):
    if l.startswith("{"):
        d = json.loads(l); print("vqe", d["value"], d["ms_per_step"], d["config"].get("energy_mean"), d["config"].get("grad_norm"), d.get("roofline"))
PY
tail -3 gpurun_out/bench_vqe_r2b.err
