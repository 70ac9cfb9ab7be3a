"""
ORACLE (test infrastructure, never imported by the product package).

Minimal numpy restatement of the `tensornetwork-ng==0.5.0` surface that the
TensorCircuit-NG hot path touches (SURVEY.md §8c "required shim surface").
`tensornetwork` is a third-party dependency of the reference
(/root/reference/pyproject.toml:23-29, pinned ==0.5.0 in
requirements/requirements-2411.txt) and is absent from this image, so its
published semantics are restated here [UPSTREAM-UNVERIFIED] and anchored on the
reference's own call sites:

  * tn.Node / Edge / `^`            tensorcircuit/basecircuit.py:277-293
  * tn.copy(nodes, conjugate)       tensorcircuit/basecircuit.py:161, :384
  * tn.contract_between             tensorcircuit/cons.py:396,413,450,948
  * tn.contract_parallel            tensorcircuit/cons.py:346
  * tn.get_all_edges / get_subgraph_dangling   tensorcircuit/cons.py:781,792,879
  * tn.CopyNode                     tensorcircuit/basecircuit.py:343
  * `_stable_id_` creation counter  tensorcircuit/cons.py:28-53

The in-repo precedent for the pairwise replay convention (result axes =
kept(left) ++ kept(right)) is examples/omeco_ready_wave_benchmark.py:198-263.
"""

from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

_NODE_CREATION_COUNTER = 0


def _next_id() -> int:
    global _NODE_CREATION_COUNTER
    v = _NODE_CREATION_COUNTER
    _NODE_CREATION_COUNTER += 1
    return v


class Edge:
    __slots__ = ("node1", "axis1", "node2", "axis2", "name", "is_disabled")

    def __init__(self, node1, axis1, node2=None, axis2=None, name=None):
        self.node1 = node1
        self.axis1 = axis1
        self.node2 = node2
        self.axis2 = axis2
        self.name = name
        self.is_disabled = False

    @property
    def dimension(self) -> int:
        return int(self.node1.shape[self.axis1])

    def is_dangling(self) -> bool:
        return self.node2 is None

    def is_trace(self) -> bool:
        return self.node1 is self.node2

    def get_nodes(self):
        return [self.node1, self.node2]

    def disable(self) -> None:
        self.is_disabled = True

    def __xor__(self, other: "Edge") -> "Edge":
        return connect(self, other)

    def disconnect(self) -> Tuple["Edge", "Edge"]:
        """Break a standard edge into two dangling edges (simplify.py:258-260)."""
        if self.is_dangling():
            raise ValueError("Cannot break a dangling edge.")
        n1, a1, n2, a2 = self.node1, self.axis1, self.node2, self.axis2
        e1 = Edge(n1, a1)
        e2 = Edge(n2, a2)
        n1.add_edge(e1, a1, override=True)
        n2.add_edge(e2, a2, override=True)
        self.disable()
        return e1, e2

    def __repr__(self) -> str:
        if self.is_dangling():
            return f"Edge(Dangling)[{self.axis1}]"
        return f"Edge({self.node1.name}[{self.axis1}]->{self.node2.name}[{self.axis2}])"


class Node:
    def __init__(self, tensor, name: Optional[str] = None, axis_names=None, backend=None):
        self.tensor = tensor if not isinstance(tensor, Node) else tensor.tensor
        self.name = name if name is not None else "__unnamed_node__"
        self.edges: List[Edge] = [Edge(self, i) for i in range(len(self.tensor.shape))]
        self.backend = backend
        self._stable_id_ = _next_id()

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.tensor.shape)

    def get_rank(self) -> int:
        return len(self.tensor.shape)

    def __getitem__(self, i: int) -> Edge:
        return self.edges[i]

    def get_edge(self, i: int) -> Edge:
        return self.edges[i]

    def __iter__(self):
        return iter(self.edges)

    def get_all_edges(self) -> List[Edge]:
        return list(self.edges)

    def get_all_dangling(self) -> List[Edge]:
        return [e for e in self.edges if e.is_dangling()]

    def get_all_nondangling(self) -> Set[Edge]:
        return {e for e in self.edges if not e.is_dangling()}

    def add_edge(self, edge: Edge, axis: int, override: bool = False) -> None:
        self.edges[axis] = edge

    def reorder_edges(self, edge_order: Sequence[Edge]) -> "Node":
        if set(map(id, edge_order)) != set(map(id, self.edges)) or len(edge_order) != len(self.edges):
            raise ValueError("Given edge order does not match expected edges.")
        perm = []
        for e in edge_order:
            for i, mine in enumerate(self.edges):
                if mine is e and i not in perm:
                    perm.append(i)
                    break
        self.tensor = np.transpose(self.tensor, perm) if len(perm) else self.tensor
        self.edges = list(edge_order)
        seen: Dict[int, int] = {}
        for i, e in enumerate(self.edges):
            # a trace edge appears twice: first occurrence -> axis1, second -> axis2
            if e.node1 is self and e.node2 is self:
                if id(e) in seen:
                    e.axis2 = i
                else:
                    e.axis1 = i
                    seen[id(e)] = i
            elif e.node1 is self:
                e.axis1 = i
            else:
                e.axis2 = i
        return self

    def copy(self, conjugate: bool = False) -> "Node":
        t = np.conj(self.tensor) if conjugate else self.tensor
        n = self.__class__.__new__(self.__class__)
        Node.__init__(n, t, name=self.name)
        return n

    def __repr__(self) -> str:
        return f"Node({self.name}, shape={self.shape}, id={self._stable_id_})"


class CopyNode(Node):
    """Generalised delta tensor (hyperedge). tensorcircuit/basecircuit.py:343."""

    def __init__(self, rank: int, dimension: int, name: Optional[str] = None, dtype=np.complex64):
        t = np.zeros((dimension,) * rank, dtype=dtype)
        for i in range(dimension):
            t[(i,) * rank] = 1
        super().__init__(t, name=name)
        self.rank = rank
        self.dimension = dimension

    def copy(self, conjugate: bool = False) -> "CopyNode":
        return CopyNode(self.rank, self.dimension, name=self.name, dtype=self.tensor.dtype)


def connect(edge1: Edge, edge2: Edge, name: Optional[str] = None) -> Edge:
    for e in (edge1, edge2):
        if not e.is_dangling():
            raise ValueError(f"Edge '{e}' is not a dangling edge.")
    if edge1 is edge2:
        raise ValueError("Cannot connect an edge to itself.")
    if edge1.dimension != edge2.dimension:
        raise ValueError("Cannot connect edges of unequal dimension.")
    n1, a1 = edge1.node1, edge1.axis1
    n2, a2 = edge2.node1, edge2.axis1
    new_edge = Edge(n1, a1, n2, a2, name=name)
    n1.add_edge(new_edge, a1, override=True)
    n2.add_edge(new_edge, a2, override=True)
    edge1.disable()
    edge2.disable()
    return new_edge


def get_all_edges(nodes: Iterable[Node]) -> Set[Edge]:
    edges: Set[Edge] = set()
    for n in nodes:
        edges |= set(n.edges)
    return edges


def get_shared_edges(node1: Node, node2: Node) -> Set[Edge]:
    nodes = {id(node1), id(node2)}
    return {
        e
        for e in node1.edges
        if (not e.is_dangling()) and {id(e.node1), id(e.node2)} == nodes
    }


def get_subgraph_dangling(nodes: Iterable[Node]) -> Set[Edge]:
    ns = set(map(id, nodes))
    out: Set[Edge] = set()
    for n in nodes:
        for e in n.edges:
            if e.is_dangling() or id(e.node1) not in ns or id(e.node2) not in ns:
                out.add(e)
    return out


def copy(nodes: Iterable[Node], conjugate: bool = False) -> Tuple[Dict[Node, Node], Dict[Edge, Edge]]:
    nodes = list(nodes)
    node_dict: Dict[Node, Node] = {}
    for n in nodes:
        node_dict[n] = n.copy(conjugate)
    edge_dict: Dict[Edge, Edge] = {}
    done: Set[int] = set()
    for n in nodes:
        for e in n.edges:
            if id(e) in done:
                continue
            done.add(id(e))
            n1, a1 = e.node1, e.axis1
            if e.is_dangling() or e.node2 not in node_dict or n1 not in node_dict:
                # keep it dangling on whichever endpoint was copied
                if n1 in node_dict:
                    ne = Edge(node_dict[n1], a1, name=e.name)
                    node_dict[n1].add_edge(ne, a1)
                else:
                    ne = Edge(node_dict[e.node2], e.axis2, name=e.name)
                    node_dict[e.node2].add_edge(ne, e.axis2)
                edge_dict[e] = ne
                continue
            ne = Edge(node_dict[n1], a1, node_dict[e.node2], e.axis2, name=e.name)
            node_dict[n1].add_edge(ne, a1)
            node_dict[e.node2].add_edge(ne, e.axis2)
            edge_dict[e] = ne
    return node_dict, edge_dict


def contract_between(
    node1: Node,
    node2: Node,
    name: Optional[str] = None,
    allow_outer_product: bool = False,
) -> Node:
    """tensordot over all shared edges; result edges = remaining(node1) ++ remaining(node2)."""
    if node1 is node2:
        return _contract_trace_edges(node1, name)
    shared = get_shared_edges(node1, node2)
    if not shared and not allow_outer_product:
        raise ValueError(f"No edges found between nodes '{node1}' and '{node2}'")
    axes1: List[int] = []
    axes2: List[int] = []
    for i, e in enumerate(node1.edges):
        if e in shared:
            axes1.append(i)
            axes2.append(e.axis2 if e.node1 is node1 else e.axis1)
    new_tensor = np.tensordot(node1.tensor, node2.tensor, [axes1, axes2])
    new_node = Node(new_tensor, name=name)
    kept: List[Tuple[Edge, Node, int]] = []
    for i, e in enumerate(node1.edges):
        if e not in shared:
            kept.append((e, node1, i))
    for i, e in enumerate(node2.edges):
        if e not in shared:
            kept.append((e, node2, i))
    _attach(new_node, kept)
    for e in shared:
        e.disable()
    node1.edges = []
    node2.edges = []
    return new_node


def _attach(new_node: Node, kept: List[Tuple[Edge, Node, int]]) -> None:
    """Point the surviving edges of the parent(s) at `new_node` (axis = position)."""
    new_node.edges = [k[0] for k in kept]
    first_seen: Dict[int, int] = {}
    for i, (e, parent, ax) in enumerate(kept):
        if e.node1 is parent and e.node2 is parent:  # trace edge carried over
            if id(e) in first_seen:
                j = first_seen[id(e)]
                e.node1, e.axis1, e.node2, e.axis2 = new_node, j, new_node, i
            else:
                first_seen[id(e)] = i
            continue
        if e.node1 is parent and e.axis1 == ax:
            e.node1, e.axis1 = new_node, i
        else:
            e.node2, e.axis2 = new_node, i


def _contract_trace_edges(node: Node, name: Optional[str] = None) -> Node:
    t = node.tensor
    edges = list(node.edges)
    while True:
        tr = next((e for e in edges if e.node1 is node and e.node2 is node), None)
        if tr is None:
            break
        i, j = [k for k, e in enumerate(edges) if e is tr]
        t = np.trace(t, axis1=i, axis2=j)
        edges = [e for k, e in enumerate(edges) if k not in (i, j)]
        tr.disable()
        # renumber remaining edge axes relative to `node`
        for k, e in enumerate(edges):
            if e.node1 is node and e.node2 is node:
                ks = [m for m, x in enumerate(edges) if x is e]
                e.axis1, e.axis2 = ks[0], ks[1]
            elif e.node1 is node:
                e.axis1 = k
            else:
                e.axis2 = k
    new_node = Node(t, name=name)
    kept = []
    for k, e in enumerate(edges):
        kept.append((e, node, k))
    _attach(new_node, kept)
    node.edges = []
    return new_node


def contract_parallel(edge: Edge) -> Node:
    """tensorcircuit/cons.py:346 — contract every edge parallel to `edge`."""
    if edge.is_dangling():
        raise ValueError("Attempted to contract dangling edge")
    return contract_between(edge.node1, edge.node2, allow_outer_product=False)


def outer_product(node1: Node, node2: Node, name: Optional[str] = None) -> Node:
    return contract_between(node1, node2, name=name, allow_outer_product=True)
