"""Developer micro-benchmark (not the driver's bench.py): time the fused evolution of the QAOA workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import tensorcircuit_ng_b200 as tc
from tensorcircuit_ng_b200 import svengine, _lib, passplan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.set_default_device("cuda:0")
import networkx as nx
g = nx.random_regular_graph(3, n, seed=0)
rng = np.random.default_rng(0)
gam, bet = rng.uniform(0, np.pi, p), rng.uniform(0, np.pi, p)

def circ():
    c = tc.Circuit(n)
    for q in range(n): c.h(q)
    for l in range(p):
        for a, b in g.edges: c.exp1(int(a), int(b), unitary=tc.gates._zz_matrix, theta=float(gam[l]))
        for q in g.nodes: c.rx(int(q), theta=float(bet[l]))
    return c

t0 = time.time(); c = circ(); t1 = time.time()
psi = c.wavefunction(); torch.cuda.synchronize(); t2 = time.time()
print(f"build {t1-t0:.3f}s first wavefunction {t2-t1:.3f}s norm {float(torch.linalg.vector_norm(psi)):.6f}")
del psi
# steady state: device time of the evolution only
nodes, d_edges = c._copy()
nq, init, gates = svengine.extract_gate_stream(nodes, d_edges)
structure = [(gg[1], svengine.gate_kind(gg[0], gg[2]), int(gg[0].tensor.numel())) for gg in gates]
cc = svengine.compile_circuit(n, structure, torch.device("cuda:0"), absorb_prefix=True)
gatebuf = svengine.build_gatebuf([gg[0].tensor for gg in gates], torch.device("cuda:0"))
plan = cc.plan
print("gates", plan.n_gates, "passes", plan.n_passes, "launches", plan.n_launches)
state = svengine.new_zero_state(n, 1, torch.device("cuda:0"))
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cc.start_and_run(state, gatebuf); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    byt = plan.n_passes * 16 * 2**n
    print(f"iter {it}: {ms:.2f} ms  {len(gates)/ms*1e3:.0f} gates/s  pass-GB/s {byt/ms/1e6:.0f}  per-pass {ms/plan.n_passes:.3f} ms")
# per-pass timing
pi = 0
for step in plan.steps:
    if isinstance(step, passplan.PassStep):
        prog_ptr = cc.programs.data_ptr() + 4 * cc.offsets[pi]; pi += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("tcb_sv_run_pass", state.data_ptr(), n, 1, prog_ptr, len(step.program), step.tile_bits, step.low_bits, step.pool_elems, gatebuf.data_ptr(), 0, 0, _lib.stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"  pass {pi}: gates {len(step.gate_ids)} subpasses {step.n_subpasses} {ms:.3f} ms  {16*2**n/ms/1e6:.0f} GB/s")
# expectation timings
psi = state
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); zz = tc.expect.z_expectations(psi, n, [[int(a), int(b)] for a, b in g.edges]); e1.record(); torch.cuda.synchronize()
print(f"batched z_expectations ({len(g.edges)} terms): {e0.elapsed_time(e1):.3f} ms  {8*2**n/e0.elapsed_time(e1)/1e6:.0f} GB/s")
e0.record(); v = tc.expect.pauli_expectation(psi, n, [], [], [0, 1]); e1.record(); torch.cuda.synchronize()
print(f"single ZZ: {e0.elapsed_time(e1):.3f} ms; x-type:", end=" ")
e0.record(); v = tc.expect.pauli_expectation(psi, n, [3], [], []); e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1):.3f} ms")
