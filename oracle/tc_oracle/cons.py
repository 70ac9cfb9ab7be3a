"""
ORACLE (test infrastructure, never imported by the product package).

numpy restatement of the contractor surface of TensorCircuit-NG
(/root/reference/tensorcircuit/cons.py).  Each function cites the lines it
follows.  Planner code (opt_einsum) is restated in `paths.py`.
"""

from __future__ import annotations

import sys
from collections import deque
from contextlib import contextmanager
from functools import partial, reduce
from operator import mul
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import paths, tn

thismodule = sys.modules[__name__]
dtypestr = "complex64"
rdtypestr = "float32"
npdtype = np.complex64


# cons.py:56-69 -----------------------------------------------------------------
def _get_edge_stable_key(edge: tn.Edge) -> Tuple[int, int, int, int]:
    n1, n2 = edge.node1, edge.node2
    id1 = getattr(n1, "_stable_id_", -1)
    id2 = getattr(n2, "_stable_id_", -1) if n2 is not None else -2
    if id1 > id2 or (id1 == id2 and edge.axis1 > edge.axis2):
        id1, id2, ax1, ax2 = id2, id1, edge.axis2, edge.axis1
    else:
        ax1, ax2 = edge.axis1, edge.axis2
    return (id1, ax1, id2, ax2)


def sorted_edges(edges) -> List[tn.Edge]:
    return sorted(edges, key=_get_edge_stable_key)


# cons.py:472-489 ----------------------------------------------------------------
_einsum_symbols_base = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i: int) -> str:
    if i < 52:
        return _einsum_symbols_base[i]
    i += 140
    if i >= 55296:
        i += 2048
    return chr(i)


def _multi_remove(elems: List[Any], indices: List[int]) -> List[Any]:  # simplify.py:83-85
    s = set(indices)
    return [e for i, e in enumerate(elems) if i not in s]


def _sizen(node: tn.Node, is_log: bool = False) -> int:  # cons.py:291-295
    s = reduce(mul, tuple(node.tensor.shape) + (1,))
    if is_log:
        return int(np.log2(s))
    return s


# cons.py:298-374 ----------------------------------------------------------------
def _merge_single_gates(nodes: List[Any], total_size: Optional[int] = None) -> Tuple[List[Any], int]:
    nodes = list(nodes)
    if total_size is None:
        total_size = sum(_sizen(t) for t in nodes)
    node_pos: Dict[int, int] = {id(n): i for i, n in enumerate(nodes)}
    queue = deque(n for n in nodes if len(n.tensor.shape) <= 2)
    in_queue: Set[int] = {id(n) for n in queue}
    while queue:
        while queue and id(queue[0]) not in in_queue:
            queue.popleft()
        if not queue:
            break
        n0 = queue[0]
        in_queue.discard(id(n0))
        try:
            n0[0]
        except IndexError:
            continue
        if n0[0].is_dangling():
            try:
                e0 = n0[1]
            except IndexError:
                continue
            if e0.is_dangling():
                continue
        else:
            e0 = n0[0]
        ep1, ep2 = e0.node1, e0.node2
        id1, id2 = id(ep1), id(ep2)
        i1 = node_pos[id1]
        i2 = node_pos[id2]
        new_node = tn.contract_parallel(e0)
        total_size += _sizen(new_node)
        in_queue.discard(id1)
        in_queue.discard(id2)
        if i1 != i2:
            early, late = (i1, i2) if i1 < i2 else (i2, i1)
            del node_pos[id1]
            del node_pos[id2]
            nodes[early] = None
            nodes[late] = new_node
            node_pos[id(new_node)] = late
        else:
            del node_pos[id1]
            nodes[i1] = new_node
            node_pos[id(new_node)] = i1
        if len(new_node.tensor.shape) <= 2:
            queue.appendleft(new_node)
            in_queue.add(id(new_node))
    nodes = [n for n in nodes if n is not None]
    return nodes, total_size


# cons.py:429-463 ----------------------------------------------------------------
def plain_contractor(nodes, output_edge_order=None, ignore_edge_order: bool = False) -> Any:
    nodes = list(reversed(list(nodes)))
    while len(nodes) > 1:
        new_node = tn.contract_between(nodes[-1], nodes[-2], allow_outer_product=True)
        nodes = _multi_remove(nodes, [len(nodes) - 2, len(nodes) - 1])
        nodes.append(new_node)
    final_node = nodes[0]
    if output_edge_order is not None:
        final_node.reorder_edges(output_edge_order)
    return final_node


# cons.py:773-800 ----------------------------------------------------------------
def _get_path_cache_friendly(nodes: List[tn.Node], algorithm: Any):
    nodes = list(nodes)
    nodes_new = sorted(nodes, key=lambda node: getattr(node, "_stable_id_", -1))
    all_edges = tn.get_all_edges(nodes_new)
    all_edges_sorted = sorted_edges(all_edges)
    mapping_dict: Dict[int, str] = {}
    i = 0
    for edge in all_edges_sorted:
        if id(edge) not in mapping_dict:
            mapping_dict[id(edge)] = get_symbol(i)
            i += 1
    input_sets = [list([mapping_dict[id(e)] for e in node.edges]) for node in nodes_new]
    output_set = list([mapping_dict[id(e)] for e in sorted_edges(tn.get_subgraph_dangling(nodes_new))])
    size_dict = {mapping_dict[id(edge)]: edge.dimension for edge in all_edges_sorted}
    return algorithm(input_sets, output_set, size_dict), nodes_new


def _identity(*args: Any, **kws: Any) -> Any:
    return args


get_tn_info = partial(_get_path_cache_friendly, algorithm=_identity)


class UnionFind:
    def __init__(self) -> None:
        self.p: Dict[int, Any] = {}
        self.obj: Dict[int, Any] = {}

    def __getitem__(self, x: Any) -> Any:
        k = id(x)
        if k not in self.p:
            self.p[k] = k
            self.obj[k] = x
        r = k
        while self.p[r] != r:
            r = self.p[r]
        while self.p[k] != r:
            self.p[k], k = r, self.p[k]
        return self.obj[r]

    def union(self, a: Any, b: Any) -> None:
        ra, rb = id(self[a]), id(self[b])
        if ra != rb:
            self.p[rb] = ra


# cons.py:492-547 ----------------------------------------------------------------
def _extract_topology(nodes: List[tn.Node]):
    nodes = sorted(nodes, key=lambda node: getattr(node, "_stable_id_", -1))
    regular_nodes = [n for n in nodes if not isinstance(n, tn.CopyNode)]
    copy_nodes = [n for n in nodes if isinstance(n, tn.CopyNode)]
    uf = UnionFind()
    for edge in tn.get_all_edges(nodes):
        uf[edge]
    for cn in copy_nodes:
        edges = cn.edges
        if edges:
            root_edge = edges[0]
            for i in range(1, len(edges)):
                uf.union(root_edge, edges[i])
    mapping_dict: Dict[int, str] = {}
    roots: Dict[int, Any] = {}
    symbol_counter = 0
    input_sets = []
    raw_tensors = []
    for node in regular_nodes:
        node_symbols = []
        for edge in node.edges:
            root = uf[edge]
            if id(root) not in mapping_dict:
                mapping_dict[id(root)] = get_symbol(symbol_counter)
                roots[id(root)] = root
                symbol_counter += 1
            node_symbols.append(mapping_dict[id(root)])
        input_sets.append("".join(node_symbols))
        raw_tensors.append(node.tensor)
    dangling_edges = sorted_edges(tn.get_subgraph_dangling(nodes))
    output_set = []
    for edge in dangling_edges:
        root = uf[edge]
        if id(root) not in mapping_dict:
            mapping_dict[id(root)] = get_symbol(symbol_counter)
            roots[id(root)] = root
            symbol_counter += 1
        output_set.append(mapping_dict[id(root)])
    size_dict = {sym: roots[k].dimension for k, sym in mapping_dict.items()}
    return raw_tensors, input_sets, "".join(output_set), size_dict


def contract_path_einsum(raw_tensors, input_sets, output_set, path) -> np.ndarray:
    """Execute a linear path with pairwise einsum (hyper-indices allowed).
    Stands in for cotengra's `make_contractor(tree, implementation="autoray")`
    (cons.py:737-738) [UPSTREAM-UNVERIFIED]: same pairwise order, kept indices =
    those still needed by the output or by another remaining operand."""
    tensors = list(raw_tensors)
    terms = [list(s) for s in input_sets]
    out = list(output_set)
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        rest = set(out)
        for k, t in enumerate(terms):
            if k not in (i, j):
                rest |= set(t)
        k12 = [s for s in dict.fromkeys(a + b) if s in rest]
        sym = {s: _einsum_symbols_base[n] for n, s in enumerate(dict.fromkeys(a + b))}
        expr = "".join(sym[s] for s in a) + "," + "".join(sym[s] for s in b) + "->" + "".join(sym[s] for s in k12)
        r = np.einsum(expr, tensors[i], tensors[j])
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
            tensors.pop(k)
        terms.append(k12)
        tensors.append(r)
    assert len(tensors) == 1
    final, t = terms[0], tensors[0]
    sym = {s: _einsum_symbols_base[n] for n, s in enumerate(dict.fromkeys(final + out))}
    return np.einsum("".join(sym[s] for s in final) + "->" + "".join(sym[s] for s in out), t)


# cons.py:706-766 ----------------------------------------------------------------
def _algebraic_base_contraction(nodes, algorithm, output_edge_order=None, ignore_edge_order=False, **kws):
    raw_tensors, input_sets, output_set, size_dict = _extract_topology(nodes)
    if len(raw_tensors) == 1:
        final_raw_tensor = contract_path_einsum(raw_tensors, input_sets, output_set, [])
    else:
        path = algorithm(input_sets, output_set, size_dict)
        final_raw_tensor = contract_path_einsum(raw_tensors, input_sets, output_set, path)
    final_node = tn.Node(final_raw_tensor)
    dangling_edges = sorted_edges(tn.get_subgraph_dangling(nodes))
    nodes_set = set(map(id, nodes))
    for i, edge in enumerate(dangling_edges):
        if id(edge.node1) in nodes_set:
            edge.node1 = final_node
            edge.axis1 = i
        else:
            edge.node2 = final_node
            edge.axis2 = i
    final_node.edges = list(dangling_edges)
    if not ignore_edge_order:
        if output_edge_order is None:
            output_edge_order = dangling_edges
        final_node.reorder_edges(list(output_edge_order))
    return final_node


# cons.py:845-961 ----------------------------------------------------------------
def _base(
    nodes,
    algorithm,
    output_edge_order=None,
    ignore_edge_order: bool = False,
    total_size: Optional[int] = None,
    debug_level: int = 0,
    use_primitives: Optional[bool] = None,
    **kws: Any,
):
    nodes_set = set(nodes)
    edges = tn.get_all_edges(nodes_set)
    if not ignore_edge_order:
        if output_edge_order is None:
            output_edge_order = list(tn.get_subgraph_dangling(nodes))
            if len(output_edge_order) > 1:
                raise ValueError(
                    "The final node after contraction has more than "
                    "one remaining edge. In this case `output_edge_order` "
                    "has to be provided."
                )
        if set(output_edge_order) != tn.get_subgraph_dangling(nodes):
            raise ValueError(
                "output edges are not equal to the remaining "
                "non-contracted edges of the final node."
            )
    has_hyperedges = any(isinstance(n, tn.CopyNode) for n in nodes)
    if use_primitives is True or (use_primitives is None and has_hyperedges):
        return _algebraic_base_contraction(nodes, algorithm, output_edge_order, ignore_edge_order, **kws)

    for edge in edges:
        if not edge.is_disabled:
            if edge.is_trace():
                idx = [i for i, n in enumerate(nodes) if id(n) == id(edge.node1)]
                nodes = _multi_remove(list(nodes), idx)
                nodes.append(tn.contract_parallel(edge))
    if len(nodes) == 1:
        if ignore_edge_order:
            return list(nodes)[0]
        return list(nodes)[0].reorder_edges(output_edge_order)

    path, nodes = _get_path_cache_friendly(nodes, algorithm)
    if debug_level == 2:
        shape = [e.dimension for e in output_edge_order] if output_edge_order else []
        return tn.Node(np.zeros(shape, dtype=npdtype))
    for ab in path:
        if len(ab) < 2:
            continue
        a, b = ab
        new_node = tn.contract_between(nodes[a], nodes[b], allow_outer_product=True)
        nodes.append(new_node)
        nodes = _multi_remove(nodes, [a, b])
    final_node = nodes[0]
    if not ignore_edge_order:
        final_node.reorder_edges(output_edge_order)
    return final_node


# cons.py:1007-1050 --------------------------------------------------------------
def custom(
    nodes,
    optimizer,
    memory_limit=None,
    output_edge_order=None,
    ignore_edge_order: bool = False,
    debug_level: int = 0,
    use_primitives: Optional[bool] = None,
    **kws: Any,
):
    local_kws = dict(kws)
    debug_level = local_kws.pop("debug_level", debug_level)
    if len(nodes) < 5:
        return _base(
            nodes,
            paths.optimal,
            output_edge_order,
            ignore_edge_order,
            debug_level=debug_level,
            use_primitives=use_primitives,
            **local_kws,
        )
    total_size = None
    has_hyperedges = any(isinstance(n, tn.CopyNode) for n in nodes)
    if local_kws.get("preprocessing", None) and not has_hyperedges:
        nodes, total_size = _merge_single_gates(nodes)
    if not isinstance(optimizer, list):
        alg = partial(optimizer, memory_limit=memory_limit)
    else:
        alg = optimizer
    return _base(
        nodes,
        alg,
        output_edge_order,
        ignore_edge_order,
        total_size,
        debug_level=debug_level,
        use_primitives=use_primitives,
        **local_kws,
    )


class NodesReturn(Exception):  # cons.py:964-973
    def __init__(self, value_to_return: Any):
        self.value = value_to_return
        super().__init__("Intentionally stopping execution to return nodes")


def _get_sorted_nodes(nodes: List[Any], *args: Any, **kws: Any) -> Any:  # cons.py:976-978
    nodes_new = sorted(nodes, key=lambda node: getattr(node, "_stable_id_", -1))
    raise NodesReturn(nodes_new)


contractor: Callable[..., Any]


def _set_global_contractor(cf: Callable[..., Any]) -> None:  # cons.py:84-87
    for module in list(sys.modules):
        if module.startswith("tc_oracle"):
            setattr(sys.modules[module], "contractor", cf)


# cons.py:1123-1261 --------------------------------------------------------------
def set_contractor(
    method: Optional[str] = None,
    optimizer: Optional[Any] = None,
    memory_limit: Optional[int] = None,
    set_global: bool = True,
    debug_level: int = 0,
    use_primitives: Optional[bool] = None,
    **kws: Any,
) -> Callable[..., Any]:
    if not method:
        method = "greedy"
    if method == "plain":
        cf: Callable[..., Any] = plain_contractor
    elif method == "before":
        cf = _get_sorted_nodes
    else:
        if method != "custom":
            optimizer = getattr(paths, method)
        cf = partial(
            custom,
            optimizer=optimizer,
            memory_limit=memory_limit,
            debug_level=debug_level,
            use_primitives=use_primitives,
            **kws,
        )
    if set_global:
        _set_global_contractor(cf)
    return cf


@contextmanager
def runtime_contractor(*confargs: Any, **confkws: Any) -> Iterator[Any]:  # cons.py:1297-1314
    old = getattr(thismodule, "contractor")
    nc = set_contractor(*confargs, **confkws)
    try:
        yield nc
    finally:
        _set_global_contractor(old)


@contextmanager
def runtime_nodes_capture(key: str = "nodes") -> Iterator[Any]:  # cons.py:994-1004
    old = getattr(thismodule, "contractor")
    set_contractor(method="before")
    captured: Dict[str, List[tn.Node]] = {}
    try:
        yield captured
    except NodesReturn as e:
        captured[key] = e.value
    finally:
        _set_global_contractor(old)


contractor = set_contractor("greedy", preprocessing=True, set_global=False)  # cons.py:1264
