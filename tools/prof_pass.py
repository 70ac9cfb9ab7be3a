"""Run a few fused passes at n qubits so ncu can capture the pass kernel (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, networkx as nx
import tensorcircuit_ng_b200 as tc
from tensorcircuit_ng_b200 import svengine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.set_default_device("cuda:0")
g = nx.random_regular_graph(3, n, seed=0)
c = tc.Circuit(n)
for q in range(n): c.h(q)
for l in range(p):
    for a, b in g.edges: c.exp1(int(a), int(b), unitary=tc.gates._zz_matrix, theta=0.3 + l)
    for q in g.nodes: c.rx(int(q), theta=0.2 + l)
psi = c.wavefunction(); torch.cuda.synchronize()
psi = c.wavefunction(); torch.cuda.synchronize()
print("ok", float(torch.linalg.vector_norm(psi)))
