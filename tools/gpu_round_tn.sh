#!/bin/bash
# ncu evidence for the tensor-network kernels: one fat GEMM-shaped step on the tensor-core kernel, one skinny
# step on the streaming kernel.  usage: tools/gpu_round_tn.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
TCB_TN_KERNEL=tc ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f \
    -o gpurun_out/prof_tc_fat_$TAG python tools/tn_prof_case.py 13 11 11 > gpurun_out/ncu_tc_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_contract -s 1 -c 1 -f \
    -o gpurun_out/prof_stream_$TAG python tools/tn_prof_case.py 26 1 1 > gpurun_out/ncu_stream_$TAG.log 2>&1
python tools/tn_bench.py > gpurun_out/tn_bench_$TAG.log 2>&1; cat gpurun_out/tn_bench_$TAG.log
python - <<'PY' 2>&1 | tee gpurun_out/tf32_peak_$TAG.log
import torch
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
for tf32 in (True, False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.matmul(a, b); torch.cuda.synchronize(); best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"torch.matmul fp32 8192^3 allow_tf32={tf32}: {best:.3f} ms = {2*8192**3/best/1e9:.1f} TFLOP/s")
PY
