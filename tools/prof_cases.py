"""Developer tool: run selected synthetic pass programs for ncu (n qubits, case name)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorcircuit_ng_b200 import passplan, svengine
from tensorcircuit_ng_b200.passplan import GateOp
n = int(sys.argv[1]); case = sys.argv[2]
dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
def rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2); return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex64)
def zz(t): return np.diag(np.exp(-1j * t * np.array([1, -1, -1, 1]))).astype(np.complex64)
hi = list(range(9)); lo = [n - 1, n - 2, n - 3, n - 4]
cases = {
    "rx12": [([q], ("dense",), rx(0.3 + q)) for q in hi[:8] + lo],
    "rx13": [([q], ("dense",), rx(0.3 + q)) for q in hi + lo],
    "rx5": [([q], ("dense",), rx(0.3 + q)) for q in hi[:5]],
    "zz45": [([int(a), int(b)], ("diag",), zz(0.2)) for a, b in (rng.permutation(n)[:2] for _ in range(45))],
    "rx1": [([0], ("dense",), rx(0.3))],
}
gates = cases[case]
ops, bufs, off = [], [], 0
for qubits, kind, mat in gates:
    ops.append(GateOp(tuple(qubits), kind, off)); bufs.append(mat.reshape(-1)); off += mat.size
plan = passplan.compile_plan(ops, n)
cc = svengine.CompiledCircuit(plan, ops, dev)
gatebuf = torch.from_numpy(np.concatenate(bufs)).to(dev)
state = svengine.new_zero_state(n, 1, dev)
for _ in range(3): cc.run(state, gatebuf)
torch.cuda.synchronize(); print("done", case, plan.n_passes)
