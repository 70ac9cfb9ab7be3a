"""
Graph layer the contractor boundary is written against.

The reference builds its networks from `tensornetwork` objects
(tensorcircuit/basecircuit.py:277-293, tensorcircuit/cons.py:742-761).  That
package is not installed in this image, so this module provides the same
duck-typed surface (`Node`, `Edge`, `CopyNode`, `^`, `copy`, `contract_between`,
`get_all_edges`, `get_subgraph_dangling`, `_stable_id_`) over torch tensors.
The engine (`cons.b200_contractor`) only relies on that surface, so real
`tensornetwork` nodes produced by an installed TensorCircuit-NG are accepted
unchanged.

Unlike the reference, pairwise contraction is NOT `backend.tensordot`: every
numeric contraction issued here goes to the CUDA kernel behind
`tcb_tn_contract` (tnengine.contract_pair); the bra copy is a lazy conjugate
view, never materialised.
"""

from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import torch

_NODE_CREATION_COUNTER = 0


def _next_id() -> int:
    global _NODE_CREATION_COUNTER
    v = _NODE_CREATION_COUNTER
    _NODE_CREATION_COUNTER += 1
    return v


class Edge:
    __slots__ = ("node1", "axis1", "node2", "axis2", "name", "is_disabled")

    def __init__(self, node1: "Node", axis1: int, node2: Optional["Node"] = None,
                 axis2: Optional[int] = None, name: Optional[str] = None) -> None:  # fmt: skip
        self.node1, self.axis1, self.node2, self.axis2 = node1, axis1, node2, axis2
        self.name = name
        self.is_disabled = False

    @property
    def dimension(self) -> int:
        return int(self.node1.shape[self.axis1])

    def is_dangling(self) -> bool:
        return self.node2 is None

    def is_trace(self) -> bool:
        return self.node1 is self.node2

    def disable(self) -> None:
        self.is_disabled = True

    def other(self, node: "Node", axis: int) -> Tuple[Optional["Node"], Optional[int]]:
        """The endpoint that is not (node, axis)."""
        if self.node1 is node and self.axis1 == axis:
            return self.node2, self.axis2
        return self.node1, self.axis1

    def __xor__(self, other: "Edge") -> "Edge":
        return connect(self, other)

    def disconnect(self) -> Tuple["Edge", "Edge"]:
        if self.is_dangling():
            raise ValueError("Cannot break a dangling edge.")
        n1, a1, n2, a2 = self.node1, self.axis1, self.node2, self.axis2
        e1, e2 = Edge(n1, a1), Edge(n2, a2)
        n1.edges[a1] = e1
        n2.edges[a2] = e2
        self.disable()
        return e1, e2


class Node:
    def __init__(self, tensor: Any, name: Optional[str] = None, axis_names: Any = None,
                 backend: Any = None) -> None:  # fmt: skip
        if isinstance(tensor, Node):
            tensor = tensor.tensor
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor)
        self.tensor = tensor
        self.name = name if name is not None else "__unnamed_node__"
        self.edges: List[Edge] = [Edge(self, i) for i in range(tensor.dim())]
        self.backend = backend
        self._stable_id_ = _next_id()

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.tensor.shape)

    def get_rank(self) -> int:
        return self.tensor.dim()

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

    def add_edge(self, edge: Edge, axis: int, override: bool = False) -> None:
        self.edges[axis] = edge

    def reorder_edges(self, edge_order: Sequence[Edge]) -> "Node":
        if len(edge_order) != len(self.edges) or {id(e) for e in edge_order} != {id(e) for e in self.edges}:
            raise ValueError("Given edge order does not match expected edges.")
        perm: List[int] = []
        for e in edge_order:
            for i, mine in enumerate(self.edges):
                if mine is e and i not in perm:
                    perm.append(i)
                    break
        if perm != list(range(len(perm))):
            self.tensor = self.tensor.permute(perm)
        self.edges = list(edge_order)
        seen: Set[int] = set()
        for i, e in enumerate(self.edges):
            if e.node1 is self and e.node2 is self:
                if id(e) in seen:
                    e.axis2 = i
                else:
                    e.axis1 = i
                    seen.add(id(e))
            elif e.node1 is self:
                e.axis1 = i
            else:
                e.axis2 = i
        return self

    def copy(self, conjugate: bool = False) -> "Node":
        # lazy conjugate view (torch conj bit): the bra is never materialised
        t = self.tensor.conj() if conjugate else self.tensor
        n = self.__class__.__new__(self.__class__)
        Node.__init__(n, t, name=self.name)
        for attr in ("_b200_kind",):
            if hasattr(self, attr):
                setattr(n, attr, getattr(self, attr))
        return n

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.name!r}, shape={self.shape}, id={self._stable_id_})"


class CopyNode(Node):
    """Generalised delta (hyperedge), tensorcircuit/basecircuit.py:343."""

    def __init__(self, rank: int, dimension: int, name: Optional[str] = None,
                 dtype: Any = torch.complex64, device: Any = None) -> None:  # fmt: skip
        self.rank = rank
        self.dimension = dimension
        self._dtype = dtype
        self._device = device
        self._tensor: Optional[torch.Tensor] = None
        self.name = name if name is not None else "__unnamed_node__"
        self.edges = []
        self.backend = None
        self._stable_id_ = _next_id()
        self.edges = [Edge(self, i) for i in range(rank)]

    @property
    def shape(self) -> Tuple[int, ...]:  # type: ignore[override]
        return (self.dimension,) * self.rank

    def get_rank(self) -> int:
        return self.rank

    @property
    def tensor(self) -> torch.Tensor:  # type: ignore[override]
        if self._tensor is None:
            t = torch.zeros((self.dimension,) * self.rank, dtype=self._dtype, device=self._device)
            for i in range(self.dimension):
                t[(i,) * self.rank] = 1
            self._tensor = t
        return self._tensor

    @tensor.setter
    def tensor(self, v: torch.Tensor) -> None:
        self._tensor = v

    def copy(self, conjugate: bool = False) -> "CopyNode":
        return CopyNode(self.rank, self.dimension, name=self.name, dtype=self._dtype, device=self._device)


def connect(edge1: Edge, edge2: Edge, name: Optional[str] = None) -> Edge:
    for e in (edge1, edge2):
        if not e.is_dangling():
            raise ValueError(f"Edge '{e}' is not a dangling edge.")
    if edge1 is edge2:
        raise ValueError("Cannot connect an edge to itself.")
    if edge1.dimension != edge2.dimension:
        raise ValueError("Cannot connect edges of unequal dimension.")
    n1, a1, n2, a2 = edge1.node1, edge1.axis1, edge2.node1, edge2.axis1
    new_edge = Edge(n1, a1, n2, a2, name=name)
    n1.edges[a1] = new_edge
    n2.edges[a2] = new_edge
    edge1.disable()
    edge2.disable()
    return new_edge


def get_all_edges(nodes: Iterable[Node]) -> Set[Edge]:
    out: Set[Edge] = set()
    for n in nodes:
        out.update(n.edges)
    return out


def get_shared_edges(node1: Node, node2: Node) -> Set[Edge]:
    want = {id(node1), id(node2)}
    return {e for e in node1.edges if not e.is_dangling() and {id(e.node1), id(e.node2)} == want}


def get_subgraph_dangling(nodes: Iterable[Node]) -> Set[Edge]:
    nodes = list(nodes)
    ids = {id(n) for n in nodes}
    out: Set[Edge] = set()
    for n in nodes:
        for e in n.edges:
            if e.is_dangling() or id(e.node1) not in ids or id(e.node2) not in ids:
                out.add(e)
    return out


def copy(nodes: Iterable[Node], conjugate: bool = False) -> Tuple[Dict[Node, Node], Dict[Edge, Edge]]:
    nodes = list(nodes)
    node_dict: Dict[Node, Node] = {n: n.copy(conjugate) for n in nodes}
    edge_dict: Dict[Edge, Edge] = {}
    for n in nodes:
        for e in n.edges:
            if e in edge_dict:
                continue
            in1 = e.node1 in node_dict
            in2 = (not e.is_dangling()) and e.node2 in node_dict
            if in1 and in2:
                ne = Edge(node_dict[e.node1], e.axis1, node_dict[e.node2], e.axis2, name=e.name)
                node_dict[e.node1].edges[e.axis1] = ne
                node_dict[e.node2].edges[e.axis2] = ne
            elif in1:
                ne = Edge(node_dict[e.node1], e.axis1, name=e.name)
                node_dict[e.node1].edges[e.axis1] = ne
            else:
                ne = Edge(node_dict[e.node2], e.axis2, name=e.name)
                node_dict[e.node2].edges[e.axis2] = ne
            edge_dict[e] = ne
    return node_dict, edge_dict


def _attach(new_node: Node, kept: List[Tuple[Edge, Node, int]]) -> None:
    new_node.edges = [k[0] for k in kept]
    first: Dict[int, int] = {}
    for i, (e, parent, ax) in enumerate(kept):
        if e.node1 is parent and e.node2 is parent:
            if id(e) in first:
                j = first[id(e)]
                e.node1, e.axis1, e.node2, e.axis2 = new_node, j, new_node, i
            else:
                first[id(e)] = i
            continue
        if e.node1 is parent and e.axis1 == ax:
            e.node1, e.axis1 = new_node, i
        else:
            e.node2, e.axis2 = new_node, i


def contract_between(node1: Node, node2: Node, name: Optional[str] = None,
                     allow_outer_product: bool = False) -> Node:  # fmt: skip
    """Contract all shared edges; result axes = remaining(node1) ++ remaining(node2)
    (same convention as tensornetwork / examples/omeco_ready_wave_benchmark.py:198-263)."""
    from . import tnengine

    if node1 is node2:
        return _contract_trace(node1, name)
    shared = get_shared_edges(node1, node2)
    if not shared and not allow_outer_product:
        raise ValueError(f"No edges found between nodes '{node1}' and '{node2}'")
    axes1: List[int] = []
    axes2: List[int] = []
    for i, e in enumerate(node1.edges):
        if e in shared:
            axes1.append(i)
            axes2.append(e.axis2 if e.node1 is node1 else e.axis1)
    new_tensor = tnengine.tensordot(node1.tensor, node2.tensor, axes1, axes2)
    new_node = Node(new_tensor, name=name)
    kept = [(e, node1, i) for i, e in enumerate(node1.edges) if e not in shared]
    kept += [(e, node2, i) for i, e in enumerate(node2.edges) if e not in shared]
    _attach(new_node, kept)
    for e in shared:
        e.disable()
    node1.edges = []
    node2.edges = []
    return new_node


def _contract_trace(node: Node, name: Optional[str] = None) -> Node:
    from . import tnengine

    t = node.tensor
    edges = list(node.edges)
    while True:
        tr = next((e for e in edges if e.node1 is node and e.node2 is node), None)
        if tr is None:
            break
        i, j = [k for k, e in enumerate(edges) if e is tr]
        t = tnengine.trace(t, i, j)
        edges = [e for k, e in enumerate(edges) if k not in (i, j)]
        tr.disable()
        for k, e in enumerate(edges):
            if e.node1 is node and e.node2 is node:
                ks = [m for m, x in enumerate(edges) if x is e]
                e.axis1, e.axis2 = ks[0], ks[1]
            elif e.node1 is node:
                e.axis1 = k
            else:
                e.axis2 = k
    new_node = Node(t, name=name)
    _attach(new_node, [(e, node, k) for k, e in enumerate(edges)])
    node.edges = []
    return new_node


def contract_parallel(edge: Edge) -> Node:
    if edge.is_dangling():
        raise ValueError("Attempted to contract dangling edge")
    return contract_between(edge.node1, edge.node2)
