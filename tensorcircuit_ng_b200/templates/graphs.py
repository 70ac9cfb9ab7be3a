"""Lattice graphs used to state Hamiltonians (`tensorcircuit/templates/graphs.py`)."""

from __future__ import annotations

from typing import Any


def Line1D(n: int, node_weight: Any = None, edge_weight: Any = None, pbc: bool = True) -> Any:
    """`tensorcircuit/templates/graphs.py` Line1D: chain of n nodes, closed into a ring when `pbc`."""
    import networkx as nx

    g = nx.Graph()
    ew = 1.0 if edge_weight is None else edge_weight
    nw = 0.0 if node_weight is None else node_weight
    ew = list(ew) if isinstance(ew, (list, tuple)) else [ew] * n
    nw = list(nw) if isinstance(nw, (list, tuple)) else [nw] * n
    for i in range(n):
        g.add_node(i, weight=nw[i])
    for i in range(n if pbc else n - 1):
        g.add_edge(i, (i + 1) % n, weight=ew[i])
    return g
