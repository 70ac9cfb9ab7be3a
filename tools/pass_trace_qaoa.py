"""Developer tool: per-phase timeline of the CTAs on SM 0 for one pass of the 30-qubit QAOA plan
(needs the -DPASS_PROFILE build, libtcb200_prof.so; CTA-per-tile kernel only: TCB_PASS_KERNEL=0)."""
import os, sys, ctypes, shutil
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
os.environ["TCB_PASS_KERNEL"] = "0"
lib = os.path.join(root, "tensorcircuit_ng_b200", "lib")
shutil.copy(os.path.join(lib, "libtcb200_prof.so"), os.path.join(lib, "libtcb200.so"))
import numpy as np, torch, networkx as nx
import tensorcircuit_ng_b200 as tc
from tensorcircuit_ng_b200 import _lib, passplan, svengine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
which = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = 8
torch.set_default_device("cuda:0")
g = nx.random_regular_graph(3, n, seed=0)
rng = np.random.default_rng(0)
gam, bet = rng.uniform(0, np.pi, p), rng.uniform(0, np.pi, p)
c = tc.Circuit(n)
for q in range(n): c.h(q)
for l in range(p):
    for a, b in g.edges: c.exp1(int(a), int(b), unitary=tc.gates._zz_matrix, theta=float(gam[l]))
    for q in g.nodes: c.rx(int(q), theta=float(bet[l]))
nodes, d_edges = c._copy()
nq, init, gates = svengine.extract_gate_stream(nodes, d_edges)
structure = [(gg[1], svengine.gate_kind(gg[0], gg[2]), int(gg[0].tensor.numel())) for gg in gates]
dev = torch.device("cuda:0")
cc = svengine.compile_circuit(n, structure, dev, absorb_prefix=True)
gatebuf = svengine.build_gatebuf([gg[0].tensor for gg in gates], dev)
state = svengine.new_zero_state(n, 1, dev)
cc.start(state, gatebuf)
L = _lib.load()
fn = L.tcb_debug_pass_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
buf = (ctypes.c_ulonglong * (8 * 64 * 6))(); ns = ctypes.c_int(0)
steps = [s for s in cc.plan.steps if isinstance(s, passplan.PassStep)]
for pi, step in enumerate(steps):
    prog_ptr = cc.programs.data_ptr() + 4 * cc.offsets[pi]
    if pi == which:
        torch.cuda.synchronize(); fn(buf, ctypes.byref(ns))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("tcb_sv_run_pass", state.data_ptr(), n, 1, prog_ptr, len(step.program), step.tile_bits, step.low_bits, step.pool_elems, gatebuf.data_ptr(), 0, 0, _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    if pi == which:
        fn(buf, ctypes.byref(ns))
        print(f"pass {pi}: gates {len(step.gate_ids)} subpasses {step.n_subpasses} words {len(step.program)} {e0.elapsed_time(e1):.3f} ms")
        break
tr = np.frombuffer(buf, dtype=np.uint64).reshape(8, 64, 6).astype(np.int64)
print("CTAs traced on SM 0:", ns.value)
t0 = tr[:ns.value, 0, 0].min()
names = ["start", "fills", "loaded", "computed", "stored", "synced"]
tot = np.zeros(5)
cnt = 0
for k in range(4, 40):
    for s in range(min(ns.value, 8)):
        r = (tr[s, k] - t0) / 1000.0
        d = np.diff(r)
        tot += d; cnt += 1
        if k < 8:
            print(f"tile {k} cta {s}: start {r[0]:8.2f}  issue+fills {d[0]:5.2f}  wait-load {d[1]:5.2f}  compute {d[2]:5.2f}  store {d[3]:5.2f}  sync {d[4]:5.2f} us   next-start {(tr[s,k+1,0]-t0)/1000.0 - r[5]:5.2f}")
print("mean per tile (us): issue+fills %.2f  wait-load %.2f  compute %.2f  store %.2f  sync %.2f  total %.2f" % (*(tot / cnt), tot.sum() / cnt))
