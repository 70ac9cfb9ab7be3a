"""cProfile of one value_and_grad sample of the config-2 TFIM VQE energy (host-side overhead hunt)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import tensorcircuit_ng_b200 as tc  # noqa: E402

n, depth = int(sys.argv[1]) if len(sys.argv) > 1 else 24, 6
dev = torch.device("cuda", 0)
torch.set_default_device(dev)


ls, ws = [], []
for q in range(n - 1):
    t = [0] * n
    t[q] = t[q + 1] = 3
    ls.append(t)
    ws.append(-1.0)
for q in range(n):
    t = [0] * n
    t[q] = 1
    ls.append(t)
    ws.append(-1.0)
ham = tc.quantum.PauliStringSum2COO(ls, ws)


def energy(p):
    c = tc.Circuit(n)
    for q in range(n):
        c.h(q)
    for l in range(depth):
        for q in range(n - 1):
            c.rzz(q, q + 1, theta=p[l, 0, q])
        for q in range(n):
            c.rx(q, theta=p[l, 1, q])
    return tc.templates.measurements.operator_expectation(c, ham)


vag = tc.backend.value_and_grad(energy)
p = 0.1 * torch.randn(depth, 2, n)
for _ in range(2):
    vag(p)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    v, g = vag(p)
torch.cuda.synchronize()
print("ms/sample", (time.perf_counter() - t0) / 3 * 1e3, float(v), float(g.norm()))
# forward only
t0 = time.perf_counter()
for _ in range(3):
    with torch.no_grad():
        e = energy(p)
torch.cuda.synchronize()
print("forward-only ms", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile()
pr.enable()
v, g = vag(p)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
