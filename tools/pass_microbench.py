"""Developer tool: time the pass kernel on synthetic programs to decompose its cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorcircuit_ng_b200 import _lib, passplan, svengine
from tensorcircuit_ng_b200.passplan import GateOp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def rx(theta):
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex64)

def zz(theta):
    return np.diag(np.exp(-1j * theta * np.array([1, -1, -1, 1]))).astype(np.complex64)

def run(label, gates, reps=3, **opts):
    ops, bufs, off = [], [], 0
    for qubits, kind, mat in gates:
        ops.append(GateOp(tuple(qubits), kind, off)); bufs.append(mat.reshape(-1)); off += mat.size
    plan = passplan.compile_plan(ops, n, **opts)
    cc = svengine.CompiledCircuit(plan, ops, dev)
    gatebuf = torch.from_numpy(np.concatenate(bufs)).to(dev) if bufs else torch.zeros(4, dtype=torch.complex64, device=dev)
    state = svengine.new_zero_state(n, 1, dev)
    cc.run(state, gatebuf); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cc.run(state, gatebuf); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    nsub = sum(s.n_subpasses for s in plan.steps if isinstance(s, passplan.PassStep))
    print(f"{label:44s} passes {plan.n_passes:2d} subpasses {nsub:2d} gates {len(gates):3d}  {best:8.3f} ms  ({16*2**n*plan.n_passes/best/1e6:6.0f} GB/s)")

hi = [0, 1, 2, 3, 4, 5, 6, 7, 8]            # qubits 0..8 = high bits
lo = [n - 1, n - 2, n - 3, n - 4]           # low bits (always in tile)
run("1 rx on low qubit", [([n - 1], ("dense",), rx(0.3))])
run("1 rx on high qubit", [([0], ("dense",), rx(0.3))])
run("5 rx (1 subpass)", [([q], ("dense",), rx(0.3 + q)) for q in hi[:5]])
run("5 rx x 4 deep, same qubits (1 subpass)", [([q], ("dense",), rx(0.3 + q + d)) for d in range(4) for q in hi[:5]])
run("10 rx (2 subpasses)", [([q], ("dense",), rx(0.3 + q)) for q in hi[:5] + lo + [8]])
run("13 rx (3 subpasses)", [([q], ("dense",), rx(0.3 + q)) for q in hi + lo])
run("1 zz diag only", [([0, 1], ("diag",), zz(0.2))])
run("45 zz diag only", [([int(a), int(b)], ("diag",), zz(0.2)) for a, b in (rng.permutation(n)[:2] for _ in range(45))])
run("45 zz + 5 rx", [([int(a), int(b)], ("diag",), zz(0.2)) for a, b in (rng.permutation(n)[:2] for _ in range(45))] + [([q], ("dense",), rx(0.3)) for q in hi[:5]])
run("45 zz + 13 rx", [([int(a), int(b)], ("diag",), zz(0.2)) for a, b in (rng.permutation(n)[:2] for _ in range(45))] + [([q], ("dense",), rx(0.3)) for q in hi + lo])
run("13 rx, T=13", [([q], ("dense",), rx(0.3 + q)) for q in hi + lo], tile_bits=13)
run("12 rx, T=12", [([q], ("dense",), rx(0.3 + q)) for q in hi[:8] + lo], tile_bits=12)
run("5 rx, L=5", [([q], ("dense",), rx(0.3 + q)) for q in hi[:5]], low_bits=5)
pass
run("1 cnot high", [([0, 1], ("ctrl", 1, 1), np.array([[1,0,0,0],[0,1,0,0],[0,0,0,1],[0,0,1,0]], dtype=np.complex64))])
run("1 dense 2q (smem subpass)", [([0, 1], ("dense",), np.kron(rx(0.3), rx(0.5)))])
# reference: plain copy bandwidth
a = torch.empty(2**n, dtype=torch.complex64, device=dev); b = torch.empty_like(a)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
print(f"torch copy_: {e0.elapsed_time(e1):.3f} ms ({16*2**n/e0.elapsed_time(e1)/1e6:.0f} GB/s)")

# ---- streaming references -----------------------------------------------------------
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
state = svengine.new_zero_state(n, 1, dev)
ms = timeit(lambda: state.mul_(1.0)); print(f"torch in-place mul_: {ms:.3f} ms ({16*2**n/ms/1e6:.0f} GB/s)")
d = torch.ones(2, dtype=torch.complex64, device=dev)
ms = timeit(lambda: _lib.call("tcb_sv_apply_diag", state.data_ptr(), n, 1, _lib.int_array([5]), 1, d.data_ptr(), 1, 0, 0, _lib.stream_ptr()))
print(f"tcb_sv_apply_diag (global RMW): {ms:.3f} ms ({16*2**n/ms/1e6:.0f} GB/s)")
m = torch.eye(2, dtype=torch.complex64, device=dev)
ms = timeit(lambda: _lib.call("tcb_sv_apply_dense", state.data_ptr(), n, 1, _lib.int_array([20]), 1, m.data_ptr(), 0, _lib.stream_ptr()))
print(f"tcb_sv_apply_dense k=1 bit 20: {ms:.3f} ms ({16*2**n/ms/1e6:.0f} GB/s)")
for T in (13, 12):
    for L in (4, 3, 2):
        for name, hi in (("contiguous", list(range(L, T))), ("scattered", [L + 1 + 2 * i for i in range(T - L)])):
            tile_pos = list(range(L)) + hi; pos_of = [n - 1 - q for q in range(n)]
            step, _ = passplan._build_pass([], [], n, pos_of, tile_pos, L, None)
            prog = torch.from_numpy(step.program).to(dev)
            gb = torch.zeros(4, dtype=torch.complex64, device=dev)
            ms = timeit(lambda: _lib.call("tcb_sv_run_pass", state.data_ptr(), n, 1, prog.data_ptr(), len(step.program), T, L, step.pool_elems, gb.data_ptr(), 0, 0, _lib.stream_ptr()))
            print(f"empty pass (load->smem->store) T={T} L={L} {name}: {ms:.3f} ms ({16*2**n/ms/1e6:.0f} GB/s)")
