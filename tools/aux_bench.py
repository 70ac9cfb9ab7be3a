"""Device-time measurements of the kernels around the pass kernel (Pauli sum, fused adjoint step, sampling),
each against its algorithmic bytes.  Output: one line per case; copy to profiles/."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import tensorcircuit_ng_b200 as tc  # noqa: E402
from tensorcircuit_ng_b200 import _lib, sampling  # noqa: E402

PEAK = 6552.0
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK))
except Exception:  # pylint: disable=broad-except
    pass


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, ms, nbytes, extra=""):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms:9.3f} ms  {gbs:8.1f} GB/s  frac {gbs / PEAK:5.2f}  {extra}", flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    psi = torch.view_as_complex(torch.randn(1 << n, 2, generator=g, device=dev))
    psi /= torch.linalg.vector_norm(psi)
    sb = (1 << n) * 8
    # --- Pauli sums
    ls, ws = [], []
    for q in range(n - 1):
        t = [0] * n
        t[q] = t[q + 1] = 3
        ls.append(t)
        ws.append(-1.0)
    hz = tc.quantum.PauliStringSum(ls, ws)
    report(f"pauli_sum value, {hz.nterms} ZZ terms (1 flip group), n={n}", timed(lambda: hz.expectation(psi)), sb)
    for q in range(n):
        t = [0] * n
        t[q] = 1
        ls.append(t)
        ws.append(-1.0)
    h = tc.quantum.PauliStringSum(ls, ws)
    ngroups = len(set(h.xmask.tolist()))
    ms = timed(lambda: h.expectation(psi))
    report(f"pauli_sum value, TFIM {h.nterms} terms ({ngroups} flip groups), n={n}", ms, sb,
           f"(gathered reads: {ngroups} x state = {ngroups * sb / ms / 1e6:.0f} GB/s through L2)")
    ms = timed(lambda: h.mvp(psi))
    report(f"pauli_sum H psi, TFIM {h.nterms} terms, n={n}", ms, 2 * sb)
    # --- fused adjoint step
    lam = psi.clone()
    p2 = psi.clone()
    grad = torch.zeros(32, dtype=torch.float64, device=dev)
    for k, bits in [(1, [0]), (1, [n // 2]), (1, [n - 1]), (2, [1, 0]), (2, [n - 1, 3]), (2, [n // 2, n // 2 + 1])]:
        u = torch.linalg.qr(torch.view_as_complex(torch.randn(1 << k, 1 << k, 2, generator=g, device=dev)))[0].contiguous()
        bp = _lib.int_array(bits)
        ms = timed(lambda: _lib.call("tcb_sv_adjoint_step", lam.data_ptr(), p2.data_ptr(), n, 1, bp, k, u.data_ptr(), 0,
                                     grad.data_ptr(), 0, _lib.stream_ptr()))  # fmt: skip
        report(f"adjoint_step k={k} bits={bits}, n={n}", ms, 4 * sb)
    del lam, p2
    # --- sampling
    ms = timed(lambda: sampling.StateSampler(psi, n))
    report(f"sample_prepare (segment CDF), n={n}", ms, sb)
    s = sampling.StateSampler(psi, n)
    for shots in (1024, 65536):
        u1 = torch.rand(shots, dtype=torch.float64, device=dev, generator=g)
        ms = timed(lambda: s.draw(u1, 0))
        print(f"{'sample mode 0 (cdf), shots=' + str(shots):58s} {ms:9.3f} ms  {shots / ms / 1e3:8.2f} Mshots/s", flush=True)
        un = torch.rand(shots, n, dtype=torch.float64, device=dev, generator=g)
        ms = timed(lambda: s.draw(un, 1))
        print(f"{'sample mode 1 (conditional walk), shots=' + str(shots):58s} {ms:9.3f} ms  {shots / ms / 1e3:8.2f} Mshots/s", flush=True)


if __name__ == "__main__":
    main()
