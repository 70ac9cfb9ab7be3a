"""
Contractor plugin surface — the drop-in boundary.

Mirrors /root/reference/tensorcircuit/cons.py: the module-global `contractor`
(rebound in every loaded module of the package, cons.py:84-87), `set_contractor`
(cons.py:1123-1261), `runtime_contractor` / `set_function_contractor` scoping
(cons.py:1269-1314), the "before" capture hack (cons.py:976-1004), and the contractor
callable contract

    cf(nodes, output_edge_order=None, ignore_edge_order=False, **kws) -> tn.Node

with the reference's two ValueErrors (cons.py:886-896) and its edge re-pointing
(cons.py:742-761).  The default contractor here is `b200_contractor`:

  circuit-shaped network  -> fused statevector passes           (svengine / passplan)
  [psi, psi*, ops] network -> reduction kernels, bra never built (K3/K4)
  anything else            -> planned pairwise GPU contractions  (tnengine)

No branch computes on the CPU.
"""

from __future__ import annotations

import sys
from collections import deque
from contextlib import contextmanager
from functools import partial, wraps
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Set, Tuple

import torch

from . import _lib, planner, svengine, tn, tnengine

package_name = "tensorcircuit_ng_b200"
thismodule = sys.modules[__name__]
dtypestr = "complex64"
rdtypestr = "float32"
idtypestr = "int32"

contractor: Callable[..., Any]


def _set_global_contractor(contractor_fn: Callable[..., Any]) -> None:  # cons.py:84-87
    for module in list(sys.modules):
        if module.startswith(package_name):
            setattr(sys.modules[module], "contractor", contractor_fn)


# cons.py:56-69 -----------------------------------------------------------------------
def _get_edge_stable_key(edge: Any) -> Tuple[int, int, int, int]:
    n1, n2 = edge.node1, edge.node2
    id1 = getattr(n1, "_stable_id_", -1)
    id2 = getattr(n2, "_stable_id_", -1) if n2 is not None else -2
    if id1 > id2 or (id1 == id2 and edge.axis1 > edge.axis2):
        id1, id2, ax1, ax2 = id2, id1, edge.axis2, edge.axis1
    else:
        ax1, ax2 = edge.axis1, edge.axis2
    return (id1, ax1, id2, ax2)


def sorted_edges(edges: Any) -> List[Any]:
    return sorted(edges, key=_get_edge_stable_key)


_einsum_symbols_base = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i: int) -> str:  # cons.py:472-489
    if i < 52:
        return _einsum_symbols_base[i]
    i += 140
    if i >= 55296:
        i += 2048
    return chr(i)


def _all_edges(nodes: Sequence[Any]) -> Set[Any]:
    out: Set[Any] = set()
    for n in nodes:
        out.update(n.edges)
    return out


def _subgraph_dangling(nodes: Sequence[Any]) -> Set[Any]:
    ids = {id(n) for n in nodes}
    out: Set[Any] = set()
    for n in nodes:
        for e in n.edges:
            if e.is_dangling() or id(e.node1) not in ids or id(e.node2) not in ids:
                out.add(e)
    return out


def _is_copynode(n: Any) -> bool:
    return type(n).__name__ == "CopyNode"


class _UnionFind:
    def __init__(self) -> None:
        self.parent: Dict[int, int] = {}
        self.obj: Dict[int, Any] = {}

    def find(self, x: Any) -> Any:
        k = id(x)
        if k not in self.parent:
            self.parent[k] = k
            self.obj[k] = x
        r = k
        while self.parent[r] != r:
            r = self.parent[r]
        while self.parent[k] != r:
            self.parent[k], k = r, self.parent[k]
        return self.obj[r]

    def union(self, a: Any, b: Any) -> None:
        ra, rb = id(self.find(a)), id(self.find(b))
        if ra != rb:
            self.parent[rb] = ra


def get_tn_info(nodes: Sequence[Any], use_primitives: Optional[bool] = None):
    """(input_sets, output_set, size_dict), sorted nodes — cons.py:773-804 for plain networks and
    cons.py:492-547 (`_extract_topology`) when CopyNode hyperedges are present: same node order
    (`_stable_id_`), same edge order (`sorted_edges`), same symbol assignment.
    `use_primitives` follows cons.py:898-908: True, or None with CopyNodes present, takes the hyperedge
    description; False keeps CopyNodes as ordinary (dense delta) nodes."""
    nodes_new = sorted(nodes, key=lambda node: getattr(node, "_stable_id_", -1))
    has_hyper = any(_is_copynode(n) for n in nodes_new) if use_primitives is None else bool(use_primitives)
    if not has_hyper:
        all_edges_sorted = sorted_edges(_all_edges(nodes_new))
        mapping: Dict[int, str] = {}
        for edge in all_edges_sorted:
            if id(edge) not in mapping:
                mapping[id(edge)] = get_symbol(len(mapping))
        input_sets = [[mapping[id(e)] for e in node.edges] for node in nodes_new]
        output_set = [mapping[id(e)] for e in sorted_edges(_subgraph_dangling(nodes_new))]
        size_dict = {mapping[id(e)]: e.dimension for e in all_edges_sorted}
        return (input_sets, output_set, size_dict), nodes_new
    regular = [n for n in nodes_new if not _is_copynode(n)]
    uf = _UnionFind()
    for e in _all_edges(nodes_new):
        uf.find(e)
    for cn in nodes_new:
        if _is_copynode(cn) and cn.edges:
            for e in cn.edges[1:]:
                uf.union(cn.edges[0], e)
    mapping = {}
    roots: Dict[int, Any] = {}
    input_sets = []
    for node in regular:
        syms = []
        for e in node.edges:
            r = uf.find(e)
            if id(r) not in mapping:
                mapping[id(r)] = get_symbol(len(mapping))
                roots[id(r)] = r
            syms.append(mapping[id(r)])
        input_sets.append(syms)
    output_set = []
    for e in sorted_edges(_subgraph_dangling(nodes_new)):
        r = uf.find(e)
        if id(r) not in mapping:
            mapping[id(r)] = get_symbol(len(mapping))
            roots[id(r)] = r
        output_set.append(mapping[id(r)])
    size_dict = {sym: roots[k].dimension for k, sym in mapping.items()}
    return (input_sets, output_set, size_dict), regular


def _validate_edge_order(nodes: Sequence[Any], output_edge_order: Optional[Sequence[Any]],
                         ignore_edge_order: bool) -> Optional[List[Any]]:  # fmt: skip
    """The reference's checks and messages, cons.py:882-896."""
    if ignore_edge_order:
        return None if output_edge_order is None else list(output_edge_order)
    dangling = _subgraph_dangling(nodes)
    if output_edge_order is None:
        output_edge_order = list(dangling)
        if len(output_edge_order) > 1:
            raise ValueError(
                "The final node after contraction has more than "
                "one remaining edge. In this case `output_edge_order` "
                "has to be provided."
            )
    if set(output_edge_order) != dangling:
        raise ValueError(
            "output edges are not equal to the remaining " "non-contracted edges of the final node."
        )
    return list(output_edge_order)


def _finalize(nodes: Sequence[Any], tensor: torch.Tensor, tensor_edges: Sequence[Any],
              output_edge_order: Optional[Sequence[Any]], ignore_edge_order: bool) -> Any:  # fmt: skip
    """Wrap the result: the returned node owns the ORIGINAL dangling Edge objects, re-pointed
    to it (cons.py:742-761; pinned by the reference's tests/test_hyperedge.py:498-527)."""
    final_node = tn.Node(tensor)
    ids = {id(n) for n in nodes}
    for i, edge in enumerate(tensor_edges):
        if id(edge.node1) in ids:
            edge.node1, edge.axis1 = final_node, i
        else:
            edge.node2, edge.axis2 = final_node, i
    final_node.edges = list(tensor_edges)
    for n in nodes:
        n.edges = []  # inputs are consumed, as after tn.contract_between
    if not ignore_edge_order and output_edge_order is not None and list(output_edge_order) != list(tensor_edges):
        final_node.reorder_edges(list(output_edge_order))
    return final_node


# ---- [psi, psi*, ops...] networks (tensorcircuit/basecircuit.py:393-447) ---------------------
def _phys(t: torch.Tensor) -> torch.Tensor:
    """The physical tensor behind a torch.vmap batched tensor (storage identity / conj bit live there)."""
    f = torch._C._functorch
    while isinstance(t, torch.Tensor) and f.is_batchedtensor(t):
        t = f.get_unwrapped(t)
    return t


def _same_storage_conj(a: torch.Tensor, b: torch.Tensor) -> bool:
    a, b = _phys(a), _phys(b)
    if a.shape != b.shape or a.is_conj() == b.is_conj():
        return False
    pa, pb = (a.conj() if a.is_conj() else a), (b.conj() if b.is_conj() else b)
    return pa.data_ptr() == pb.data_ptr() and pa.stride() == pb.stride() and pa.is_contiguous()


def _recognize_expectation(nodes: Sequence[Any]):
    """Returns (ket tensor, n, [(op node, qubit axes)]) for <psi|ops|psi>, else None."""
    if len(nodes) < 2 or _subgraph_dangling(nodes):
        return None
    if any(_is_copynode(x) for x in nodes):
        return None
    # the bra is the lazily conjugated view of the ket's storage (tn.Node.copy(conjugate=True))
    bra = next((x for x in nodes if _phys(x.tensor).is_conj()), None)
    if bra is None:
        return None
    ket = next((x for x in nodes if x is not bra and _same_storage_conj(x.tensor, bra.tensor)), None)
    if ket is None:
        return None
    n = ket.get_rank()
    if n < 1:
        return None
    ops = [x for x in nodes if x is not ket and x is not bra]
    covered: Set[int] = set()
    found = []
    for op in ops:
        r = op.get_rank()
        if r % 2 or r == 0:
            return None
        k = r // 2
        axes = []
        for j in range(k):
            eb, ek = op.edges[j], op.edges[j + k]
            nb, ab = svengine._other_end(eb, op, j)
            nk, ak = svengine._other_end(ek, op, j + k)
            if nb is not bra or nk is not ket or ab != ak or ab in covered:
                return None
            covered.add(ab)
            axes.append(ab)
        found.append((op, tuple(axes)))
    for j in range(n):
        if j in covered:
            continue
        e = ket.edges[j]
        o, ax = svengine._other_end(e, ket, j)
        if o is not bra or ax != j:
            return None
    return ket.tensor, n, found


def _pauli_of(op: Any) -> Optional[str]:
    if getattr(op, "_b200_kind", None) is None or op.get_rank() != 2:
        return None
    name = str(getattr(op, "name", ""))
    return name if name in ("x", "y", "z", "i") else None


speculate_z_moments = True


def _z_moment(ket: torch.Tensor, n: int, zq: Sequence[int]) -> Optional[torch.Tensor]:
    """<Z_i> / <Z_i Z_j> on a cached state.  An energy is a sum of many such terms, each a separate
    `expectation_ps` call in the reference's API and each a full read of the state (1.4 ms at n = 30).
    The second query on the same state computes ALL one- and two-body Z moments in ceil((n + n(n-1)/2) / 64)
    reads (`tcb_sv_expect_z`, 64 strings per read) and parks them on the state tensor object; later
    queries are lookups.  Worst case (exactly two queries): ~10 reads instead of 2."""
    from . import expect

    if not speculate_z_moments or n < 2 or n > 40 or not ket.is_cuda or _phys(ket) is not ket:
        return None
    cache = getattr(ket, "_b200_zcache", None)
    if cache is None or cache["version"] != ket._version:
        cache = {"version": ket._version, "count": 0, "tables": []}
        try:
            ket._b200_zcache = cache  # type: ignore[attr-defined]
        except Exception:  # pylint: disable=broad-except  (wrapper tensors)
            return None
    cache["count"] += 1
    key = tuple(zq)
    for index, table in cache["tables"]:
        if key in index:
            return table[index[key]]
    if cache["count"] < 2:
        return None

    def build(terms: List[List[int]]) -> torch.Tensor:
        index = {tuple(t): k for k, t in enumerate(terms)}
        table = expect.z_expectations(ket.reshape(-1), n, terms).to(torch.complex64)
        cache["tables"].append((index, table))
        return table[index[key]]

    # first table: every single qubit + the pairs the circuit itself couples (the terms of a VQE / QAOA energy
    # usually follow the circuit's connectivity; hint left on the state by Circuit._copy_state_tensor) — one or
    # two reads; a query outside it pays for the full table of all pairs
    hint = getattr(ket, "_b200_pair_hint", None)
    if not cache["tables"] and hint:
        terms = [[i] for i in range(n)] + [list(p) for p in sorted(hint) if 0 <= p[0] < p[1] < n]
        if key in {tuple(t) for t in terms} and len(terms) < n + n * (n - 1) // 4:
            return build(terms)
    return build([[i] for i in range(n)] + [[i, j] for i in range(n) for j in range(i + 1, n)])


def _expectation_value(ket: torch.Tensor, n: int, ops: Sequence[Tuple[Any, Tuple[int, ...]]]) -> Optional[torch.Tensor]:
    from . import autograd, expect

    psi = ket.reshape(-1)
    paulis = [_pauli_of(op) for op, _ in ops]
    if torch.is_grad_enabled() and any(p is None and autograd.wants_grad(op.tensor) for (op, _), p in zip(ops, paulis)):
        # a trainable operator tensor: the reduction kernels treat operators as constants, so this sandwich
        # goes to the tensor-network route, whose pairwise contractions differentiate every operand
        return None
    if all(p in ("z", "i") for p in paulis) and not (ket.requires_grad and torch.is_grad_enabled()):
        zq = sorted(ax[0] for (op, ax), p in zip(ops, paulis) if p == "z")
        if 1 <= len(zq) <= 2:
            val = _z_moment(ket, n, zq)
            if val is not None:
                return val
    if all(p is not None for p in paulis):
        xs = [ax[0] for (op, ax), p in zip(ops, paulis) if p == "x"]
        ys = [ax[0] for (op, ax), p in zip(ops, paulis) if p == "y"]
        zs = [ax[0] for (op, ax), p in zip(ops, paulis) if p == "z"]
        return expect.pauli_expectation(psi, n, xs, ys, zs)
    return expect.operator_expectation(psi, n, [(op.tensor, ax) for op, ax in ops])


# ---- the default contractor -------------------------------------------------------------------
def b200_contractor(nodes: List[Any], output_edge_order: Optional[List[Any]] = None,
                    ignore_edge_order: bool = False, optimizer: Any = None, **kws: Any) -> Any:  # fmt: skip
    nodes = list(nodes)
    order = _validate_edge_order(nodes, output_edge_order, ignore_edge_order)
    if kws.get("debug_level", 0) == 2:  # cons.py:928-933: shape only
        shape = [e.dimension for e in order] if order else []
        return tn.Node(torch.zeros(shape, dtype=torch.complex64))
    # 1. statevector route
    if order and not kws.get("force_tn", False):
        try:
            state = svengine.run_circuit_network(nodes, order)
            return _finalize(nodes, state, order, order, ignore_edge_order)
        except svengine.NotCircuitShaped:
            pass
    # 2. expectation sandwich on a cached state
    if not kws.get("force_tn", False):
        rec = _recognize_expectation(nodes)
        if rec is not None:
            val = _expectation_value(*rec)
            if val is not None:
                return _finalize(nodes, val, [], order, ignore_edge_order)
    # 3. general tensor network
    return _tn_route(nodes, order, ignore_edge_order, optimizer, **kws)


def diagonal_to_hyperedges(input_sets: List[List[str]], output_set: List[str], size_dict: Dict[str, int],
                           sorted_nodes: Sequence[Any], tensors: List[Any]):  # fmt: skip
    """Exact network simplification: a DIAGONAL k-qubit gate tensor [out.., in..] (cz, rzz, rz, cphase,
    exp1(ZZ) ... — dense rank-2k `Gate` nodes in the reference, SURVEY §7 "hard parts") does not cut
    the wires it sits on: out_j and in_j are the same index.  The node becomes its diagonal, a rank-k
    tensor on k hyper-indices, which lowers the contraction width of circuit networks a lot
    (a CZ crossing a cut costs 1 bit of bond instead of 2) — what the reference only gets from its
    explicit `diagonal/cmz` hyperedge API (tensorcircuit/basecircuit.py:318-369).
    Only nodes whose structural hint says "diag" are touched; returns rewritten copies."""
    parent: Dict[str, str] = {}

    def find(x: str) -> str:
        while parent.get(x, x) != x:
            parent[x] = parent.get(parent[x], parent[x])
            x = parent[x]
        return x

    picked = []
    out_syms = set(output_set)
    for i, node in enumerate(sorted_nodes):
        kind = getattr(node, "_b200_kind", None)
        syms = input_sets[i]
        if kind is None or kind[0] != "diag" or len(syms) % 2 or len(syms) == 0 or len(set(syms)) != len(syms):
            continue
        k = len(syms) // 2
        if any(syms[j] in out_syms and syms[j + k] in out_syms for j in range(k)):
            continue  # (an isolated diagonal gate with both legs dangling keeps its matrix form)
        picked.append((i, k))
        for j in range(k):
            ra, rb = find(syms[j]), find(syms[j + k])
            if ra != rb:
                # keep the symbol of a dangling leg as the representative so the output order survives
                if rb in out_syms:
                    ra, rb = rb, ra
                parent[rb] = ra
    if not picked:
        return input_sets, output_set, size_dict, tensors
    new_inputs = [[find(x) for x in t] for t in input_sets]
    new_tensors = list(tensors)
    for i, k in picked:
        t = tensors[i]
        new_tensors[i] = t.reshape(2**k, 2**k).diagonal().reshape([2] * k)
        new_inputs[i] = new_inputs[i][:k]
    for t in new_inputs:
        if len(set(t)) != len(t):  # a wire closed on itself through a diagonal gate: leave the network alone
            return input_sets, output_set, size_dict, tensors
    new_output = [find(x) for x in output_set]
    if len(set(new_output)) != len(new_output):
        return input_sets, output_set, size_dict, tensors
    new_size = {find(k_): v for k_, v in size_dict.items()}
    return new_inputs, new_output, new_size, new_tensors


def wire_groups(input_sets: Sequence[Sequence[str]], sorted_nodes: Sequence[Any]) -> Dict[str, str]:
    """symbol -> representative symbol of its qubit wire: leg j and leg j + k of every 2k-leg gate node
    [out.., in..] (tensorcircuit/basecircuit.py:288-290) lie on the same wire.  A planner hint."""
    parent: Dict[str, str] = {}

    def find(x: str) -> str:
        while parent.get(x, x) != x:
            parent[x] = parent.get(parent[x], parent[x])
            x = parent[x]
        return x

    for syms, node in zip(input_sets, sorted_nodes):
        flag = str(getattr(node, "flag", ""))
        if len(syms) % 2 == 0 and len(syms) >= 2 and (flag.startswith("gate") or hasattr(node, "_b200_kind")):
            k = len(syms) // 2
            for j in range(k):
                ra, rb = find(syms[j]), find(syms[j + k])
                if ra != rb:
                    parent[rb] = ra
    return {x: find(x) for t in input_sets for x in t}


def _merge_single_gates(nodes: Sequence[Any]) -> List[Any]:
    """The reference's `preprocessing=True` step (tensorcircuit/cons.py:298-374): every node of rank <= 2 is
    absorbed into the neighbour on its first connected leg, newly created small nodes first.  The merged node
    takes the LATER of the two list positions and a fresh `_stable_id_`, so the node order, the edge axes and
    hence the symbols an optimizer sees afterwards are the reference's."""
    slots: List[Any] = list(nodes)
    where: Dict[int, int] = {id(n): i for i, n in enumerate(slots)}
    todo = deque(n for n in slots if len(n.shape) <= 2)
    waiting: Set[int] = {id(n) for n in todo}
    while todo:
        small = todo.popleft()
        if id(small) not in waiting:
            continue  # already absorbed as somebody's neighbour
        waiting.discard(id(small))
        bond = next((e for e in small.edges[:2] if not e.is_dangling()), None)
        if bond is None:
            continue
        a, b = bond.node1, bond.node2
        ia, ib = where.pop(id(a)), where.pop(id(b), None)
        merged = tn.contract_parallel(bond)
        waiting.discard(id(a))
        waiting.discard(id(b))
        if ib is None or ib == ia:  # a trace leg: the node is replaced in its own slot
            slots[ia] = merged
            where[id(merged)] = ia
        else:
            slots[min(ia, ib)] = None
            slots[max(ia, ib)] = merged
            where[id(merged)] = max(ia, ib)
        if len(merged.shape) <= 2:
            todo.appendleft(merged)
            waiting.add(id(merged))
    return [n for n in slots if n is not None]


def contraction_info_decorator(algorithm: Callable[..., Any]) -> Callable[..., Any]:
    """cons.py:1084-1120: print the cost summary of the path an optimizer returns (same line format; the
    numbers come from `planner.path_stats`, the cotengra accounting of SURVEY §8d)."""
    import math
    import time

    def new_algorithm(input_sets: Any, output_set: Any, size_dict: Any, **kws: Any) -> Any:
        t0 = time.time()
        path = algorithm(input_sets, output_set, size_dict, **kws)
        dt = time.time() - t0
        st = planner.path_stats(input_sets, output_set, size_dict, [p for p in path if len(p) == 2])
        print("------ contraction cost summary ------")
        print(
            "log10[FLOPs]: ", "%.3f" % math.log10(max(st["flops"], 1.0)),
            " log2[SIZE]: ", "%.0f" % math.log2(max(st["size"], 1.0)),
            " log2[WRITE]: ", "%.3f" % math.log2(max(st["write"], 1.0)),
            " PathFindingTime: ", "%.3f" % dt,
        )  # fmt: skip
        return path

    return new_algorithm


def _strip_exponent(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mantissa, exponent) with t = mantissa * 10 ** exponent and max |mantissa| = 1: the return convention of
    cotengra's `tree.contract(..., strip_exponent=True)` that cons.py:736-740,763-764 passes on
    [UPSTREAM-UNVERIFIED base-10 convention]."""
    scale = t.detach().abs().max()
    safe = torch.where(scale > 0, scale, torch.ones_like(scale))
    return t / safe, torch.log10(safe)


def _tn_route(nodes: List[Any], order: Optional[List[Any]], ignore_edge_order: bool, optimizer: Any,
              **kws: Any) -> Any:  # fmt: skip
    (input_sets, output_set, size_dict), sorted_nodes = get_tn_info(nodes, kws.get("use_primitives"))
    tensors = [n.tensor for n in sorted_nodes]
    if kws.get("hyper_diagonal", optimizer is None):
        # (a caller-supplied path / optimizer is computed for the reference's network: leave it alone)
        input_sets, output_set, size_dict, tensors = diagonal_to_hyperedges(
            input_sets, output_set, size_dict, sorted_nodes, tensors)  # fmt: skip
    device = svengine.pick_device(tensors)
    tensors = [t if t.device == device else t.to(device) for t in tensors]
    dangling = sorted_edges(_subgraph_dangling(nodes))
    if len(tensors) == 1:
        path: List[Tuple[int, ...]] = []
    elif isinstance(optimizer, list):
        path = optimizer
    elif optimizer is not None:
        path = optimizer(input_sets, output_set, size_dict)
    else:
        path = planner.greedy_alpha(input_sets, output_set, size_dict)
    # the tree writes the caller's edge order directly (no final transpose, K2)
    if order is not None and not ignore_edge_order and len(set(output_set)) == len(output_set):
        sym_of = {id(e): s for e, s in zip(dangling, output_set)}
        want = [sym_of[id(e)] for e in order]
        result_edges = list(order)
    else:
        want = list(output_set)
        result_edges = list(dangling)
    out = tnengine.contract_tree(tensors, input_sets, want, path)
    if kws.get("strip_exponent", False):
        out, exponent = _strip_exponent(out)
        return _finalize(nodes, out, result_edges, order, ignore_edge_order), exponent
    return _finalize(nodes, out, result_edges, order, ignore_edge_order)


def plain_contractor(nodes: List[Any], output_edge_order: Optional[List[Any]] = None,
                     ignore_edge_order: bool = False) -> Any:  # fmt: skip
    """cons.py:429-463: literal sequential statevector order, node by node (GPU kernel per pair)."""
    nodes = list(reversed(list(nodes)))
    while len(nodes) > 1:
        new_node = tn.contract_between(nodes[-1], nodes[-2], allow_outer_product=True)
        nodes = nodes[:-2] + [new_node]
    final_node = nodes[0]
    if output_edge_order is not None:
        final_node.reorder_edges(output_edge_order)
    return final_node


_CUSTOM_KWS = ("preprocessing", "strip_exponent", "hyper_diagonal", "contraction_info")


def custom(nodes: List[Any], optimizer: Any, memory_limit: Optional[int] = None,
           output_edge_order: Optional[List[Any]] = None, ignore_edge_order: bool = False,
           debug_level: int = 0, use_primitives: Optional[bool] = None, **kws: Any) -> Any:  # fmt: skip
    """cons.py:1007-1050 Level-1 plug: the caller's planner, our executor.  Same behaviour as the reference:
    fewer than five nodes use the exhaustive `optimal` search (:1019-1030), `preprocessing` merges rank <= 2
    nodes first unless the network has hyperedges (:1034), the optimizer may be a literal path (:1037-1040) and
    `strip_exponent` returns `(node, exponent)` (:763-764)."""
    unknown = sorted(k for k in kws if k not in _CUSTOM_KWS)
    if unknown:
        raise TypeError(f"custom contractor: unsupported option(s) {unknown}; honoured: {list(_CUSTOM_KWS)}")
    nodes = list(nodes)
    order = _validate_edge_order(nodes, output_edge_order, ignore_edge_order)
    if debug_level == 2:
        shape = [e.dimension for e in order] if order else []
        return tn.Node(torch.zeros(shape, dtype=torch.complex64))
    has_hyper = any(_is_copynode(n) for n in nodes)
    if len(nodes) < 5:
        alg: Any = planner.optimal
    else:
        if kws.get("preprocessing") and not has_hyper:
            nodes = _merge_single_gates(nodes)
        alg = optimizer if isinstance(optimizer, list) or optimizer is None else partial(optimizer, memory_limit=memory_limit)
    if alg is None:
        alg = planner.greedy
    return _tn_route(nodes, order, ignore_edge_order, alg, use_primitives=use_primitives,
                     strip_exponent=kws.get("strip_exponent", False),
                     hyper_diagonal=kws.get("hyper_diagonal", False))  # fmt: skip


def custom_stateful(nodes: List[Any], optimizer: Any, memory_limit: Optional[int] = None,
                    opt_conf: Optional[Dict[str, Any]] = None, output_edge_order: Optional[List[Any]] = None,
                    ignore_edge_order: bool = False, use_primitives: Optional[bool] = None, **kws: Any) -> Any:  # fmt: skip
    """cons.py:1053-1081: the optimizer class is instantiated afresh for every contraction."""
    opt = optimizer(**(opt_conf or {}))
    local = dict(kws)
    debug_level = local.pop("debug_level", 0)
    if local.pop("contraction_info", None):
        opt = contraction_info_decorator(opt)
    return custom(nodes, opt, memory_limit=memory_limit, output_edge_order=output_edge_order,
                  ignore_edge_order=ignore_edge_order, debug_level=debug_level, use_primitives=use_primitives,
                  **local)  # fmt: skip


class NodesReturn(Exception):  # cons.py:964-973
    def __init__(self, value_to_return: Any):
        self.value = value_to_return
        super().__init__(f"Intentionally stopping execution to return: {value_to_return}")


def _get_sorted_nodes(nodes: List[Any], *args: Any, **kws: Any) -> Any:
    raise NodesReturn(sorted(nodes, key=lambda node: getattr(node, "_stable_id_", -1)))


def _auto_path(input_sets: Any, output_set: Any, size_dict: Any, memory_limit: Optional[int] = None) -> Any:
    """opt_einsum's "auto": exhaustive below five tensors, greedy from there on (its branch-* middle tiers are
    not restated, so between 5 and 14 tensors this is greedy where opt_einsum would still search)."""
    f = planner.optimal if len(input_sets) < 5 else planner.greedy
    return f(input_sets, output_set, size_dict, memory_limit=memory_limit)


# method names the reference forwards to `getattr(opt_einsum.paths, method)` (cons.py:1245-1246)
_OPT_EINSUM_METHODS: Dict[str, Any] = {
    "greedy": planner.greedy, "eager": planner.greedy, "opportunistic": planner.greedy,
    "optimal": planner.optimal, "auto": _auto_path, "auto-hq": _auto_path,
}  # fmt: skip
_OPT_EINSUM_UNSUPPORTED = ("branch", "branch-all", "branch-2", "branch-1", "dp", "dynamic-programming",
                           "random-greedy", "random-greedy-128")  # fmt: skip


def set_contractor(method: Optional[str] = None, optimizer: Optional[Any] = None,
                   memory_limit: Optional[int] = None, opt_conf: Optional[Dict[str, Any]] = None,
                   set_global: bool = True, contraction_info: bool = False, debug_level: int = 0,
                   use_primitives: Optional[bool] = None, **kws: Any) -> Callable[..., Any]:  # fmt: skip
    """Same signature as the reference's `set_contractor` (cons.py:1123-1261).

    method:
      "b200" (default)   statevector passes / reduction kernels, planned tensor-network executor otherwise;
      "greedy", "eager", "opportunistic", "optimal", "auto", "auto-hq"
                         the opt_einsum path finders the reference resolves these names to (cons.py:1245-1246),
                         restated in `planner` (same paths as the reference's default contractor), executed by
                         the tensor-network engine through `custom`;
      "tn"               always the tensor-network engine, with this package's own planner;
      "plain"            literal node order (cons.py:429-463);
      "custom" / "custom_stateful"   the caller's `optimizer` (callable
                         `f(inputs, output, size_dict, memory_limit=None) -> path`, a literal path, or a class
                         instantiated per call with `opt_conf`);
      "before"           node capture (cons.py:976-1004).
    Options honoured like the reference: `preprocessing` (merge rank <= 2 nodes first), `contraction_info`,
    `debug_level`, `use_primitives`, `strip_exponent` (implies `use_primitives`, cons.py:1161-1163).  Anything
    else raises `TypeError` instead of being dropped.  "cotengra*" / "omeco*" need those packages and raise
    ImportError like the reference when they are absent (cons.py:678-683); opt_einsum's branch-and-bound /
    dynamic-programming / random-greedy finders are not restated and raise ValueError."""
    if not method:
        method = "b200"
    if kws.get("strip_exponent", False) and use_primitives is None:
        use_primitives = True
    if method.startswith("cotengra") or method.startswith("omeco"):
        raise ImportError(
            f"contractor {method!r} needs the optional third-party planner, which is not installed; "
            "load a plan with experimental.DistributedContractor.from_path or pass "
            "method='custom', optimizer=<callable>"
        )
    if method in _OPT_EINSUM_UNSUPPORTED:
        raise ValueError(
            f"contractor {method!r} is an opt_einsum path finder that this package does not restate (opt_einsum is "
            f"not installed); available: {sorted(_OPT_EINSUM_METHODS)} or method='custom', optimizer=<callable>"
        )
    allowed = {"b200": ("hyper_diagonal", "force_tn"), "tn": ("hyper_diagonal",), "plain": (), "before": ()}
    if method in allowed:
        bad = sorted(k for k in kws if k not in allowed[method])
        if opt_conf is not None or optimizer is not None or memory_limit is not None:
            bad.append("optimizer / opt_conf / memory_limit")
        if method in ("plain", "before") and (contraction_info or use_primitives is not None):
            bad.append("contraction_info / use_primitives")
        if bad:
            raise TypeError(f"set_contractor({method!r}): unsupported option(s) {bad}")
    if method == "plain":
        cf: Callable[..., Any] = plain_contractor
    elif method == "before":
        cf = _get_sorted_nodes
    elif method == "b200":
        cf = partial(b200_contractor, debug_level=debug_level, use_primitives=use_primitives, **kws)
    elif method == "tn":
        opt = contraction_info_decorator(planner.greedy_alpha) if contraction_info else None
        cf = partial(b200_contractor, force_tn=True, optimizer=opt, debug_level=debug_level,
                     use_primitives=use_primitives, **kws)  # fmt: skip
    elif method == "custom_stateful":
        cf = partial(custom_stateful, optimizer=optimizer, opt_conf=opt_conf, memory_limit=memory_limit,
                     contraction_info=contraction_info, debug_level=debug_level, use_primitives=use_primitives,
                     **kws)  # fmt: skip
    elif method == "custom" or method in _OPT_EINSUM_METHODS:
        if method != "custom":
            optimizer = _OPT_EINSUM_METHODS[method]
        if contraction_info is True and not isinstance(optimizer, list) and optimizer is not None:
            optimizer = contraction_info_decorator(optimizer)
        cf = partial(custom, optimizer=optimizer, memory_limit=memory_limit, debug_level=debug_level,
                     use_primitives=use_primitives, **kws)  # fmt: skip
    else:
        raise ValueError(f"Unknown contractor type: {method}")
    if set_global:
        _set_global_contractor(cf)
    return cf


get_contractor = partial(set_contractor, set_global=False)


def set_function_contractor(*confargs: Any, **confkws: Any) -> Callable[..., Any]:  # cons.py:1269-1294
    def wrapper(f: Callable[..., Any]) -> Callable[..., Any]:
        @wraps(f)
        def newf(*args: Any, **kws: Any) -> Any:
            old = getattr(thismodule, "contractor")
            set_contractor(*confargs, **confkws)
            try:
                return f(*args, **kws)
            finally:
                _set_global_contractor(old)

        return newf

    return wrapper


@contextmanager
def runtime_contractor(*confargs: Any, **confkws: Any) -> Iterator[Any]:  # cons.py:1297-1314
    old = getattr(thismodule, "contractor")
    nc = set_contractor(*confargs, **confkws)
    try:
        yield nc
    finally:
        _set_global_contractor(old)


def function_nodes_capture(func: Callable[..., Any]) -> Callable[..., Any]:  # cons.py:981-991
    @wraps(func)
    def wrapper(*args: Any, **kwargs: Any) -> Any:
        with runtime_contractor(method="before"):
            try:
                return func(*args, **kwargs)
            except NodesReturn as e:
                return e.value

    return wrapper


@contextmanager
def runtime_nodes_capture(key: str = "nodes") -> Iterator[Any]:  # cons.py:994-1004
    old = getattr(thismodule, "contractor")
    set_contractor(method="before")
    captured: Dict[str, List[Any]] = {}
    try:
        yield captured
    except NodesReturn as e:
        captured[key] = e.value
    finally:
        _set_global_contractor(old)


contractor = set_contractor("b200", set_global=False)
