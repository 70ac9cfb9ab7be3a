"""Developer tool: pairwise-contraction kernels on GEMM-shaped steps (TFLOP/s = 8 M N K / t)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorcircuit_ng_b200 import tnengine

def bench(nm, nn, nk, nb=0, reps=5, planned=False):
    letters = [chr(ord("a") + i) for i in range(nb + nm + nn + nk)]
    bat, ms, ns, ks = letters[:nb], letters[nb:nb + nm], letters[nb + nm:nb + nm + nn], letters[nb + nm + nn:]
    ma, mb, mc = bat + ms + ks, bat + ks + ns, bat + ms + ns   # the layouts tensordot would produce
    if planned:  # what tnengine's layout planning hands the kernel: contracted modes lowest in BOTH operands
        mb = bat + ns + ks
    a = torch.randn([2] * len(ma), dtype=torch.complex64, device="cuda")
    b = torch.randn([2] * len(mb), dtype=torch.complex64, device="cuda")
    out = torch.empty([2] * len(mc), dtype=torch.complex64, device="cuda")
    def run(): tnengine.contract_raw(a, ma, b, mb, mc, out=out)
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    # torch reference: the reference's own GPU path (tensordot -> cgemm)
    A2 = a.reshape(2**nb, 2**nm, 2**nk)
    B2 = b.reshape(2**nb, 2**nn, 2**nk).transpose(1, 2) if planned else b.reshape(2**nb, 2**nk, 2**nn)
    torch.bmm(A2, B2); torch.cuda.synchronize()
    tb = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.bmm(A2, B2); e1.record(); torch.cuda.synchronize(); tb = min(tb, e0.elapsed_time(e1))
    flops = 8.0 * 2 ** (nb + nm + nn + nk)
    byt = 8.0 * (2 ** (nb + nm + nk) + 2 ** (nb + nk + nn) + 2 ** (nb + nm + nn))
    print(f"{'planned ' if planned else ''}M=2^{nm} N=2^{nn} K=2^{nk} b=2^{nb}: ours {best:8.3f} ms {flops/best/1e9:8.1f} TFLOP/s {byt/best/1e6:7.0f} GB/s | torch.bmm {tb:8.3f} ms {flops/tb/1e9:8.1f} TFLOP/s", flush=True)

print("kernel:", os.environ.get("TCB_TN_KERNEL", "auto"))
for s in [(26, 1, 1), (25, 2, 1), (24, 2, 2), (22, 3, 3), (24, 4, 4), (20, 6, 6), (18, 7, 7), (16, 8, 8), (14, 10, 10), (12, 12, 12)]:
    bench(*s)
for s in [(20, 6, 6), (18, 7, 7), (16, 8, 8), (14, 10, 10), (12, 12, 12), (20, 10, 10)]:
    bench(*s, planned=True)
