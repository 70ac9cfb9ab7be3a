"""Developer tool: per-phase cycle breakdown of the pass kernel (needs the -DPASS_PROFILE build)."""
import os, sys, ctypes, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
lib = os.path.join(root, "tensorcircuit_ng_b200", "lib")
shutil.copy(os.path.join(lib, "libtcb200_prof.so"), os.path.join(lib, "libtcb200.so"))
import numpy as np, torch
from tensorcircuit_ng_b200 import _lib, passplan, svengine
from tensorcircuit_ng_b200.passplan import GateOp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0"); rng = np.random.default_rng(0)
def rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2); return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex64)
def zz(t): return np.diag(np.exp(-1j * t * np.array([1, -1, -1, 1]))).astype(np.complex64)
hi = list(range(9)); lo = [n - 1, n - 2, n - 3, n - 4]
cases = {
    "rx1": [([0], ("dense",), rx(0.3))],
    "rx4": [([q], ("dense",), rx(0.3 + q)) for q in hi[:4]],
    "rx12": [([q], ("dense",), rx(0.3 + q)) for q in hi[:8] + lo],
    "zz45": [([int(a), int(b)], ("diag",), zz(0.2)) for a, b in (rng.permutation(n)[:2] for _ in range(45))],
}
L = _lib.load()
fn = L.tcb_debug_pass_prof; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = ["issue-load", "fill", "wait+sync", "-", "subpasses", "store", "endsync"]
for case, gates in cases.items():
    ops, bufs, off = [], [], 0
    for qubits, kind, mat in gates:
        ops.append(GateOp(tuple(qubits), kind, off)); bufs.append(mat.reshape(-1)); off += mat.size
    plan = passplan.compile_plan(ops, n)
    cc = svengine.CompiledCircuit(plan, ops, dev)
    gatebuf = torch.from_numpy(np.concatenate(bufs)).to(dev)
    state = svengine.new_zero_state(n, 1, dev)
    cc.run(state, gatebuf); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)(); fn(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cc.run(state, gatebuf); e1.record(); torch.cuda.synchronize()
    fn(out, 1)
    ntiles = 2 ** (n - 12)
    tot = sum(out[i] for i in range(7))
    print(f"{case}: {e0.elapsed_time(e1):.3f} ms; cycles per tile: " + ", ".join(f"{names[i]} {out[i]/ntiles:.0f}" for i in range(7)) + f"  total {tot/ntiles:.0f}")
    print(f"    inside the register sub-passes (warp 0, cycles per tile): LDS phase {out[8]/ntiles:.0f}, rounds {out[9]/ntiles:.0f}, STS phase {out[10]/ntiles:.0f}, barrier wait {out[11]/ntiles:.0f}")
