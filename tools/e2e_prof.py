"""cProfile of one end-to-end QAOA step (public API from host parameters) — host-side overhead hunt."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import tensorcircuit_ng_b200 as tc  # noqa: E402

n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 30, 8
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.set_default_device(dev)
edges, gam, bet = bench.qaoa_problem(n, p)
zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
gam_pin, bet_pin = torch.from_numpy(gam).pin_memory(), torch.from_numpy(bet).pin_memory()


def step():
    g_d = gam_pin.to(dev, non_blocking=True)
    b_d = bet_pin.to(dev, non_blocking=True)
    t0 = time.perf_counter()
    cq = bench.build_qaoa(tc, n, edges, g_d, b_d, zz)
    t1 = time.perf_counter()
    total = None
    first = None
    for a, b in edges:
        v = cq.expectation_ps(z=[a, b])
        if first is None:
            first = time.perf_counter()
        total = v if total is None else total + v
    t2 = time.perf_counter()
    out = float((0.5 * (len(edges) - total.real)).cpu())
    t3 = time.perf_counter()
    return out, (t1 - t0, first - t1, t2 - first, t3 - t2)


for _ in range(2):
    step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    out, parts = step()
    torch.cuda.synchronize()
    print("step ms %.1f  build %.1f  first expectation (plan+launch) %.1f  other expectations %.1f  sync+readback %.1f"
          % ((time.perf_counter() - t0) * 1e3, *(x * 1e3 for x in parts)))
pr = cProfile.Profile()
pr.enable()
step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
