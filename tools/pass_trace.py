"""Developer tool: timeline of the CTAs sharing SM 0 during one pass (needs the -DPASS_PROFILE build)."""
import os, sys, ctypes, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
lib = os.path.join(root, "tensorcircuit_ng_b200", "lib")
shutil.copy(os.path.join(lib, "libtcb200_prof.so"), os.path.join(lib, "libtcb200.so"))
import numpy as np, torch
from tensorcircuit_ng_b200 import _lib, passplan, svengine
from tensorcircuit_ng_b200.passplan import GateOp
n = 30
dev = torch.device("cuda:0")
def rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2); return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex64)
hi = list(range(9)); lo = [n - 1, n - 2, n - 3, n - 4]
gates = [([q], ("dense",), rx(0.3 + q)) for q in hi[:8] + lo]
ops, bufs, off = [], [], 0
for qubits, kind, mat in gates:
    ops.append(GateOp(tuple(qubits), kind, off)); bufs.append(mat.reshape(-1)); off += mat.size
plan = passplan.compile_plan(ops, n)
cc = svengine.CompiledCircuit(plan, ops, dev)
gatebuf = torch.from_numpy(np.concatenate(bufs)).to(dev)
state = svengine.new_zero_state(n, 1, dev)
L = _lib.load()
fn = L.tcb_debug_pass_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
buf = (ctypes.c_ulonglong * (8 * 64 * 6))(); ns = ctypes.c_int(0)
cc.run(state, gatebuf); torch.cuda.synchronize(); fn(buf, ctypes.byref(ns))
cc.run(state, gatebuf); torch.cuda.synchronize(); fn(buf, ctypes.byref(ns))
tr = np.frombuffer(buf, dtype=np.uint64).reshape(8, 64, 6).astype(np.int64)
print("CTAs traced on SM 0:", ns.value, "passes", plan.n_passes)
t0 = tr[:ns.value, 0, 0].min()
names = ["start", "fills", "loaded", "computed", "stored", "synced"]
for k in range(2, 8):
    for s in range(min(ns.value, 8)):
        r = (tr[s, k] - t0) / 1000.0
        print(f"tile {k} cta {s}: " + " ".join(f"{names[i]} {r[i]:8.2f}" for i in range(6)) + f"   load {r[2]-r[0]:5.2f} compute {r[3]-r[2]:5.2f} store {r[5]-r[3]:5.2f} us")
