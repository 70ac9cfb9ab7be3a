"""
Contraction planner used when no external optimizer is supplied.

The reference delegates planning to third-party code (`opt_einsum.paths.greedy`
at tensorcircuit/cons.py:1246, cotengra's hyper-optimizer at
tensorcircuit/experimental.py:936-945), neither of which is installed here.  The plan
*interchange format* is the reference's `tree_data` dict
(tensorcircuit/experimental.py:947-953): {inputs, output, size_dict, path, sliced_inds}
— a plan produced by real cotengra elsewhere loads unchanged (experimental.py here).
Plans made by this module are labelled "ours": a deterministic greedy search
(memory-removed cost, optional Boltzmann-free restarts with a cost-weight sweep) plus a
greedy slicer that pins indices until every intermediate fits `target_size`.

Index sets are Python ints used as bitsets (fast for the 10^3-node amplitude networks of
config 5).  Paths are opt_einsum *linear* paths (pop both operands, append the result).
"""

from __future__ import annotations

import heapq
import itertools
import math
from typing import Any, Dict, List, Optional, Sequence, Tuple


def _popcount(x: int) -> int:
    return bin(x).count("1")


def _bits(m: int):
    """Indices of the set bits of a Python-int bitset (only the set ones are visited)."""
    while m:
        low = m & -m
        yield low.bit_length() - 1
        m ^= low


class _Net:
    def __init__(self, inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int]):
        syms: Dict[str, int] = {}
        for term in list(inputs) + [output]:
            for s in term:
                if s not in syms:
                    syms[s] = len(syms)
        self.syms = syms
        self.names = list(syms)
        self.log2 = [math.log2(size_dict[s]) for s in self.names]
        self.uniform2 = all(size_dict[s] == 2 for s in self.names)
        self.inputs = [self.mask(t) for t in inputs]
        self.output = self.mask(output)

    def mask(self, term: Sequence[str]) -> int:
        m = 0
        for s in term:
            m |= 1 << self.syms[s]
        return m

    def lsize(self, m: int) -> float:
        if self.uniform2:
            return float(_popcount(m))
        return sum(self.log2[i] for i in _bits(m))


def _ssa_greedy(net: _Net, alpha: float = 1.0, sliced: int = 0) -> List[Tuple[int, int]]:
    """Greedy SSA path. cost(a,b) = size(out) - alpha*(size(a)+size(b)); ties by ids."""
    terms: Dict[int, int] = {}
    for i, m in enumerate(net.inputs):
        terms[i] = m & ~sliced
    output = net.output & ~sliced
    nxt = len(terms)
    ssa: List[Tuple[int, int]] = []
    # occurrence count of every index among live terms (+1 if in the output)
    nsym = len(net.names)
    occ = [0] * nsym
    where: List[set] = [set() for _ in range(nsym)]
    for tid, m in terms.items():
        for i in _bits(m):
            occ[i] += 1
            where[i].add(tid)

    def result_of(a: int, b: int) -> int:
        ma, mb = terms[a], terms[b]
        both = ma & mb
        keep = (ma | mb) & output
        rest = (ma | mb) & ~output
        # an index survives when a tensor other than the two operands still carries it
        for i in _bits(rest & both):
            if occ[i] > 2:
                keep |= 1 << i
        for i in _bits(rest & ~both):
            if occ[i] > 1:
                keep |= 1 << i
        return keep

    def cost(a: int, b: int) -> Tuple[float, int, int]:
        r = result_of(a, b)
        c = 2.0 ** net.lsize(r) - alpha * (2.0 ** net.lsize(terms[a]) + 2.0 ** net.lsize(terms[b]))
        lo, hi = (a, b) if a < b else (b, a)
        return (c, hi, lo)

    heap: List[Tuple[Tuple[float, int, int], int, int]] = []

    def push_neighbours(t: int) -> None:
        nb = set()
        for i in _bits(terms[t]):
            nb.update(where[i])
        nb.discard(t)
        best = None
        for o in nb:
            c = cost(t, o)
            if best is None or c < best[0]:
                best = (c, t, o)
        if best is not None:
            heapq.heappush(heap, best)

    for t in list(terms):
        push_neighbours(t)

    def merge(a: int, b: int) -> int:
        nonlocal nxt
        r = result_of(a, b)
        for t in (a, b):
            for i in _bits(terms[t]):
                occ[i] -= 1
                where[i].discard(t)
            del terms[t]
        new = nxt
        nxt += 1
        terms[new] = r
        for i in _bits(r):
            occ[i] += 1
            where[i].add(new)
        ssa.append((a, b))
        return new

    while heap:
        c, a, b = heapq.heappop(heap)
        if a not in terms or b not in terms:
            if a in terms:
                push_neighbours(a)
            elif b in terms:
                push_neighbours(b)
            continue
        c2 = cost(a, b)
        if c2 != c:  # stale entry (occurrence counts changed since it was pushed): re-queue
            heapq.heappush(heap, (c2, a, b))
            continue
        new = merge(a, b)
        push_neighbours(new)
    # disconnected components: outer products, smallest first
    rest = sorted(terms, key=lambda t: (net.lsize(terms[t]), t))
    while len(rest) > 1:
        a, b = rest[0], rest[1]
        new = merge(a, b)
        rest = sorted(terms, key=lambda t: (net.lsize(terms[t]), t))
    return ssa


def ssa_to_linear(ssa: Sequence[Tuple[int, int]], n_inputs: int) -> List[Tuple[int, int]]:
    live = list(range(n_inputs))
    nxt = n_inputs
    out: List[Tuple[int, int]] = []
    for a, b in ssa:
        ia, ib = live.index(a), live.index(b)
        out.append((ia, ib))
        for k in sorted((ia, ib), reverse=True):
            live.pop(k)
        live.append(nxt)
        nxt += 1
    return out


def path_stats(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
               path: Sequence[Tuple[int, ...]], sliced: Sequence[str] = ()) -> Dict[str, float]:  # fmt: skip
    """Per-slice cost of a linear path with `sliced` indices removed.
    flops = sum_steps prod(size of every index of the pair)  (scalar complex MACs, SURVEY §8d "C"),
    write = sum of intermediate sizes, size = largest intermediate, all as plain numbers.
    Reference counting on bitsets: O(indices touched) per step."""
    net = _Net(inputs, output, size_dict)
    smask = net.mask([s for s in sliced if s in net.syms])
    terms = [m & ~smask for m in net.inputs]
    out = net.output & ~smask
    occ = [0] * len(net.names)
    for m in terms:
        for i in _bits(m):
            occ[i] += 1
    flops = write = size = 0.0
    per_step: List[Tuple[int, int, int, int]] = []
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        union, both = a | b, a & b
        keep = union & out
        for x in _bits(union & ~out):
            if occ[x] - (2 if (both >> x) & 1 else 1) > 0:
                keep |= 1 << x
        f = 2.0 ** net.lsize(union)
        w = 2.0 ** net.lsize(keep)
        flops += f
        write += w
        size = max(size, w)
        per_step.append((_popcount(a), _popcount(b), _popcount(keep), _popcount(union)))
        for x in _bits(a):
            occ[x] -= 1
        for x in _bits(b):
            occ[x] -= 1
        for x in _bits(keep):
            occ[x] += 1
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
        terms.append(keep)
    nslices = 1.0
    for s in sliced:
        nslices *= size_dict[s]
    return {"flops": flops, "write": write, "size": size, "nslices": nslices, "steps": per_step}  # type: ignore[dict-item]


def greedy_alpha(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
                 memory_limit: Optional[int] = None, alpha: float = 1.0) -> List[Tuple[int, int]]:  # fmt: skip
    """This module's own pairwise greedy (cost = out - alpha (a + b) on log sizes), Level-1 plug signature."""
    if len(inputs) == 1:
        return []
    net = _Net(inputs, output, size_dict)
    return ssa_to_linear(_ssa_greedy(net, alpha), len(inputs))


# ---- the reference's default planners -----------------------------------------------------------
# `tc.set_contractor("greedy")` resolves to `opt_einsum.paths.greedy` (tensorcircuit/cons.py:1245-1246,
# default at :1264) and `custom` uses `opt_einsum.paths.optimal` below five nodes (:1019-1030).  opt_einsum
# (pinned 3.4.0, requirements/requirements-2411.txt) is absent from this image: the two functions below restate
# its published algorithms on Python-int bitsets so that the product picks the SAME path as the reference's
# default contractor (tests/test_host_logic.py pins them to the oracle's independent set-based restatement).
def _prod_sizes(mask: int, sizes: Sequence[int]) -> int:
    r = 1
    for i in _bits(mask):
        r *= sizes[i]
    return r


def _ssa_opt_einsum_greedy(inputs: Sequence[int], output: int, sizes: Sequence[int]) -> List[Tuple[int, ...]]:
    """SSA path of opt_einsum's `ssa_greedy_optimize` with the default 'memory-removed' cost: repeatedly contract
    the pair of tensors sharing an index whose result is smallest relative to its operands
    (size(out) - size(a) - size(b)), ties by the larger, then the smaller, SSA id; identical index sets are
    merged at once; what is left at the end is combined by outer products, smallest first."""
    if len(inputs) == 1:
        return [(0,)]
    common = inputs[0]
    for m in inputs[1:]:
        common &= m
    output = output | common
    live: Dict[int, int] = {}  # index set -> SSA id
    nxt = len(inputs)
    ssa: List[Tuple[int, ...]] = []
    for i, key in enumerate(inputs):
        if key in live:  # Hadamard product of equal index sets
            ssa.append((live[key], i))
            live[key] = nxt
            nxt += 1
        else:
            live[key] = i
    holders: Dict[int, set] = {}  # contracted index -> index sets that carry it
    for key in live:
        for d in _bits(key & ~output):
            holders.setdefault(d, set()).add(key)
    ge2 = ge3 = 0  # indices carried by >= 2 / >= 3 live tensors
    for d, ks in holders.items():
        if len(ks) >= 2:
            ge2 |= 1 << d
        if len(ks) >= 3:
            ge3 |= 1 << d
    foot = {key: _prod_sizes(key, sizes) for key in live}
    heap: List[Tuple[int, int, int, int, int, int]] = []

    def push_best(k1: int, partners: Sequence[int]) -> None:
        best = None
        for k2 in partners:
            either, two = k1 | k2, k1 & k2
            one = either & ~two
            k12 = (either & output) | (two & ge3) | (one & ge2)
            cost = _prod_sizes(k12, sizes) - foot[k1] - foot[k2]
            a, b = k1, k2
            ia, ib = live[a], live[b]
            if ia > ib:
                a, ia, b, ib = b, ib, a, ia
            cand = (cost, ib, ia, a, b, k12)
            if best is None or cand[:3] < best[:3]:
                best = cand
        if best is not None:
            heapq.heappush(heap, best)

    for d, ks in holders.items():
        order = sorted(ks, key=live.__getitem__)
        for i, k1 in enumerate(order[:-1]):
            push_best(k1, order[i + 1:])
    while heap:
        _, _, _, k1, k2, k12 = heapq.heappop(heap)
        if k1 not in live or k2 not in live:
            continue
        id1, id2 = live.pop(k1), live.pop(k2)
        for d in _bits(k1 & ~output):
            holders[d].discard(k1)
        for d in _bits(k2 & ~output):
            holders[d].discard(k2)
        ssa.append((id1, id2))
        if k12 in live:
            ssa.append((live[k12], nxt))
            nxt += 1
        else:
            for d in _bits(k12 & ~output):
                holders.setdefault(d, set()).add(k12)
        live[k12] = nxt
        nxt += 1
        for d in _bits(k1 | (k2 & ~output)):
            c = len(holders.get(d, ()))
            bit = 1 << d
            ge2 = (ge2 | bit) if c >= 2 else (ge2 & ~bit)
            ge3 = (ge3 | bit) if c >= 3 else (ge3 & ~bit)
        foot[k12] = _prod_sizes(k12, sizes)
        near = set()
        for d in _bits(k12 & ~output):
            near |= holders[d]
        near.discard(k12)
        if near:
            push_best(k12, sorted(near, key=live.__getitem__))
    rest = [(_prod_sizes(key & output, sizes), i, key) for key, i in live.items()]
    heapq.heapify(rest)
    _, id1, k1 = heapq.heappop(rest)
    while rest:
        _, id2, k2 = heapq.heappop(rest)
        ssa.append((min(id1, id2), max(id1, id2)))
        k12 = (k1 | k2) & output
        _, id1, k1 = heapq.heappushpop(rest, (_prod_sizes(k12, sizes), nxt, k12))
        nxt += 1
    return ssa


def _ssa_to_linear_general(ssa: Sequence[Tuple[int, ...]]) -> List[Tuple[int, ...]]:
    n = 1 + max(max(t) for t in ssa)
    pos = list(range(n))
    out: List[Tuple[int, ...]] = []
    for ids in ssa:
        out.append(tuple(int(pos[i]) for i in ids))
        for i in ids:
            for k in range(i, n):
                pos[k] -= 1
    return out


def greedy(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
           memory_limit: Optional[int] = None) -> List[Tuple[int, ...]]:  # fmt: skip
    """The path `opt_einsum.paths.greedy` returns (the reference's default contractor,
    tensorcircuit/cons.py:1245-1246,1264), in opt_einsum's linear convention."""
    net = _Net(inputs, output, size_dict)
    sizes = [size_dict[s] for s in net.names]
    return _ssa_to_linear_general(_ssa_opt_einsum_greedy(net.inputs, net.output, sizes))


def optimal(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
            memory_limit: Optional[int] = None) -> List[Tuple[int, ...]]:  # fmt: skip
    """`opt_einsum.paths.optimal`: depth-first search over all pair orders, minimising the flop count
    (tensorcircuit/cons.py:1019-1030 uses it below five nodes)."""
    net = _Net(inputs, output, size_dict)
    sizes = [size_dict[s] for s in net.names]
    n = len(net.inputs)
    if n == 1:
        return [(0,)]
    if n > 10:
        raise ValueError(f"'optimal' enumerates all pair orders; {n} tensors is too many (use 'greedy')")
    best_flops = [float("inf")]
    best_path: List[Tuple[Tuple[int, int], ...]] = [tuple()]

    def walk(path: Tuple[Tuple[int, int], ...], alive: Tuple[int, ...], terms: Tuple[int, ...], flops: int) -> None:
        if len(alive) == 1:
            best_flops[0], best_path[0] = flops, path
            return
        for x in range(len(alive)):
            for y in range(x + 1, len(alive)):
                i, j = alive[x], alive[y]
                keep = net.output
                for r in alive:
                    if r != i and r != j:
                        keep |= terms[r]
                either = terms[i] | terms[j]
                k12 = either & keep
                f = _prod_sizes(either, sizes) * (2 if either & ~k12 else 1)
                if flops + f >= best_flops[0]:
                    continue
                rest = tuple(r for r in alive if r != i and r != j) + (len(terms),)
                walk(path + ((i, j),), rest, terms + (k12,), flops + f)

    walk(tuple(), tuple(range(n)), tuple(net.inputs), 0)
    return _ssa_to_linear_general(best_path[0])


def search(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
           target_size: Optional[int] = None, minimize: str = "flops", alphas: Sequence[float] = (1.0, 0.5, 0.75, 0.9, 1.1, 1.25, 0.25, 0.0),
           max_slices_log2: int = 40) -> Dict[str, Any]:  # fmt: skip
    """Deterministic search: greedy over a sweep of cost weights, then greedy slicing.
    Returns a `tree_data` dict (tensorcircuit/experimental.py:947-953)."""
    inputs = [tuple(t) for t in inputs]
    output = tuple(output)
    net = _Net(inputs, output, size_dict)
    best: Optional[Tuple[float, List[Tuple[int, int]], List[str]]] = None
    for alpha in alphas:
        sliced: List[str] = []
        path = ssa_to_linear(_ssa_greedy(net, alpha), len(inputs)) if len(inputs) > 1 else []
        st = path_stats(inputs, output, size_dict, path)
        if target_size is not None:
            while st["size"] > target_size and len(sliced) < max_slices_log2:
                s = _pick_slice_index(inputs, output, size_dict, path, sliced, target_size)
                if s is None:
                    break
                sliced.append(s)
                # re-plan with the index removed (cheap: greedy) and keep the better of old/new path
                mask = net.mask(sliced)
                p2 = ssa_to_linear(_ssa_greedy(net, alpha, sliced=mask), len(inputs))
                st1 = path_stats(inputs, output, size_dict, path, sliced)
                st2 = path_stats(inputs, output, size_dict, p2, sliced)
                if (st2["size"], st2["flops"]) < (st1["size"], st1["flops"]):
                    path, st = p2, st2
                else:
                    st = st1
        total = st[minimize] * st["nslices"] if minimize in ("flops", "write") else st[minimize]
        if target_size is not None and st["size"] > target_size:
            total = float("inf") if best is not None else total * 1e30
        if best is None or total < best[0]:
            best = (total, path, list(sliced))
    assert best is not None
    _, path, sliced = best
    return {
        "inputs": tuple(inputs),
        "output": output,
        "size_dict": dict(size_dict),
        "path": [tuple(p) for p in path],
        "sliced_inds": {s: size_dict[s] for s in sliced},
        "planner": "tensorcircuit_ng_b200.planner (ours; not cotengra)",
    }


def _pick_slice_index(inputs, output, size_dict, path, sliced, target_size) -> Optional[str]:
    """The index that appears in the most oversized intermediates (ties: most flops saved)."""
    net = _Net(inputs, output, size_dict)
    smask = net.mask([s for s in sliced if s in net.syms])
    terms = [m & ~smask for m in net.inputs]
    out = net.output & ~smask
    occ = [0] * len(net.names)
    for m in terms:
        for i in _bits(m):
            occ[i] += 1
    score: Dict[int, float] = {}
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        union, both = a | b, a & b
        keep = union & out
        for x in _bits(union & ~out):
            if occ[x] - (2 if (both >> x) & 1 else 1) > 0:
                keep |= 1 << x
        w = 2.0 ** net.lsize(keep)
        if w > target_size:
            f = 2.0 ** net.lsize(union)
            for x in _bits(keep & ~out):
                score[x] = score.get(x, 0.0) + w * 1e6 + f
        for x in _bits(a):
            occ[x] -= 1
        for x in _bits(b):
            occ[x] -= 1
        for x in _bits(keep):
            occ[x] += 1
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
        terms.append(keep)
    if not score:
        return None
    best = max(sorted(score, key=lambda x: net.names[x]), key=lambda x: score[x])
    return net.names[best]


def slice_values(slice_id: int, sliced_inds: Sequence[str], size_dict: Dict[str, int]) -> Dict[str, int]:
    """slice id -> value of each sliced index: mixed radix over `sliced_inds` in the given order,
    LAST index fastest (cotengra's convention as recalled in SURVEY App. C [UPSTREAM-UNVERIFIED]);
    the order used is stored in the plan so the mapping is self-describing."""
    vals: Dict[str, int] = {}
    for s in reversed(list(sliced_inds)):
        d = size_dict[s]
        vals[s] = slice_id % d
        slice_id //= d
    return vals


# ---------------------------------------------------------------------------------------------
# Recursive-bisection planner (circuit-shaped networks: 2D lattice x depth).  Greedy orderings are
# exponentially worse than a separator-based tree on such networks (7x7 depth-20 amplitude: 10^28
# vs 10^13 scalar MACs); the reference gets its trees from cotengra's hypergraph partitioner, which
# is not installed here, so this is a small stand-in: simplify (merges that do not grow a tensor),
# then recursive Kernighan-Lin bisection of the tensor graph, greedy inside small parts.
def _simplify_ssa(net: _Net, terms: Dict[int, int], output: int, nxt: int) -> Tuple[List[Tuple[int, int]], int]:
    """Merge pairs whose result is no larger than the larger operand, until none is left."""
    ssa: List[Tuple[int, int]] = []
    nsym = len(net.names)
    occ = [0] * nsym
    where: List[set] = [set() for _ in range(nsym)]
    for tid, m in terms.items():
        for i in _bits(m):
            occ[i] += 1
            where[i].add(tid)

    def result_of(ma: int, mb: int) -> int:
        union, both = ma | mb, ma & mb
        keep = union & output
        for i in _bits(union & ~output):
            if occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                keep |= 1 << i
        return keep

    work = sorted(terms)
    while work:
        t = work.pop()
        if t not in terms:
            continue
        nb = set()
        for i in _bits(terms[t]):
            nb.update(where[i])
        nb.discard(t)
        best = None
        for o in sorted(nb):
            r = result_of(terms[t], terms[o])
            if _popcount(r) <= max(_popcount(terms[t]), _popcount(terms[o])):
                key = (_popcount(r), o)
                if best is None or key < best[0]:
                    best = (key, o, r)
        if best is None:
            continue
        _, o, r = best
        for x in (t, o):
            for i in _bits(terms[x]):
                occ[i] -= 1
                where[i].discard(x)
            del terms[x]
        terms[nxt] = r
        for i in _bits(r):
            occ[i] += 1
            where[i].add(nxt)
        ssa.append((t, o))
        work.append(nxt)
        nxt += 1
    return ssa, nxt


def _bisect_tree(net: _Net, ids: List[int], masks: Dict[int, int], seed: int, leaf: int = 6) -> Any:
    """Nested tuples of term ids: recursive Kernighan-Lin bisection of the weighted tensor graph."""
    import networkx as nx
    from networkx.algorithms.community import kernighan_lin_bisection

    if len(ids) <= 1:
        return ids[0]
    if len(ids) <= leaf:
        return tuple(ids)  # small part: ordered greedily later
    g = nx.Graph()
    g.add_nodes_from(ids)
    holders: Dict[int, List[int]] = {}
    for t in ids:
        for i in _bits(masks[t]):
            holders.setdefault(i, []).append(t)
    for i, hs in holders.items():
        if len(hs) < 2:
            continue
        w = net.log2[i] / (len(hs) - 1)  # a hyper-index is one bond however many tensors hold it
        for a in range(len(hs)):
            for b in range(a + 1, len(hs)):
                if g.has_edge(hs[a], hs[b]):
                    g[hs[a]][hs[b]]["weight"] += w
                else:
                    g.add_edge(hs[a], hs[b], weight=w)
    comps = [sorted(c) for c in nx.connected_components(g)]
    if len(comps) > 1:  # disconnected: split by components (largest vs rest)
        comps.sort(key=len, reverse=True)
        a, b = comps[0], [x for c in comps[1:] for x in c]
    else:
        pa, pb = kernighan_lin_bisection(g, weight="weight", seed=seed, max_iter=20)
        a, b = sorted(pa), sorted(pb)
    return (_bisect_tree(net, a, masks, seed + 1, leaf), _bisect_tree(net, b, masks, seed + 2, leaf))


def _tree_to_ssa(net: _Net, tree: Any, terms: Dict[int, int], output: int, nxt: int) -> Tuple[List[Tuple[int, int]], int]:
    """Post-order execution of the nested tuple tree; flat tuples (leaves of the bisection) are
    ordered greedily by result size."""
    ssa: List[Tuple[int, int]] = []
    occ = [0] * len(net.names)
    for m in terms.values():
        for i in _bits(m):
            occ[i] += 1

    def merge(a: int, b: int) -> int:
        nonlocal nxt
        ma, mb = terms[a], terms[b]
        union, both = ma | mb, ma & mb
        keep = union & output
        for i in _bits(union & ~output):
            if occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                keep |= 1 << i
        for i in _bits(ma):
            occ[i] -= 1
        for i in _bits(mb):
            occ[i] -= 1
        for i in _bits(keep):
            occ[i] += 1
        del terms[a], terms[b]
        terms[nxt] = keep
        ssa.append((a, b))
        nxt += 1
        return nxt - 1

    def size_after(a: int, b: int) -> int:
        ma, mb = terms[a], terms[b]
        union, both = ma | mb, ma & mb
        n = _popcount(union & output)
        for i in _bits(union & ~output):
            if occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                n += 1
        return n

    def run(node: Any) -> int:
        if isinstance(node, int):
            return node
        if len(node) == 2 and not all(isinstance(x, int) for x in node):
            return merge(run(node[0]), run(node[1]))
        live = [run(x) for x in node]
        while len(live) > 1:
            best = None
            for x in range(len(live)):
                for y in range(x + 1, len(live)):
                    k = (size_after(live[x], live[y]), live[x], live[y])
                    if best is None or k < best[0]:
                        best = (k, x, y)
            _, x, y = best
            new = merge(live[x], live[y])
            live = [v for k, v in enumerate(live) if k not in (x, y)] + [new]
        return live[0]

    run(tree)
    return ssa, nxt


def search_bisect(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
                  target_size: Optional[int] = None, seeds: Sequence[int] = (0, 1, 2, 3), max_slices_log2: int = 40,
                  leaf: int = 6) -> Dict[str, Any]:  # fmt: skip
    """Simplify + recursive bisection + greedy slicing; returns a `tree_data` dict."""
    inputs = [tuple(t) for t in inputs]
    output = tuple(output)
    net = _Net(inputs, output, size_dict)
    best: Optional[Tuple[float, List[Tuple[int, int]], List[str]]] = None
    for seed in seeds:
        terms = {i: m for i, m in enumerate(net.inputs)}
        ssa0, nxt = _simplify_ssa(net, terms, net.output, len(inputs))
        ids = sorted(terms)
        if len(ids) > 1:
            tree = _bisect_tree(net, ids, dict(terms), seed, leaf)
            ssa1, nxt = _tree_to_ssa(net, tree, terms, net.output, nxt)
        else:
            ssa1 = []
        path = ssa_to_linear(ssa0 + ssa1, len(inputs))
        sliced: List[str] = []
        st = path_stats(inputs, output, size_dict, path)
        if target_size is not None:
            while st["size"] > target_size and len(sliced) < max_slices_log2:
                s = _pick_slice_index(inputs, output, size_dict, path, sliced, target_size)
                if s is None:
                    break
                sliced.append(s)
                st = path_stats(inputs, output, size_dict, path, sliced)
        total = st["flops"] * st["nslices"]
        if target_size is not None and st["size"] > target_size:
            total *= 1e30
        if best is None or total < best[0]:
            best = (total, path, list(sliced))
    assert best is not None
    _, path, sliced = best
    return {
        "inputs": tuple(inputs),
        "output": output,
        "size_dict": dict(size_dict),
        "path": [tuple(p) for p in path],
        "sliced_inds": {s: size_dict[s] for s in sliced},
        "planner": "tensorcircuit_ng_b200.planner.search_bisect (ours; not cotengra)",
    }


# ---------------------------------------------------------------------------------------------
# Variable-elimination planner.  Index graph (two indices are adjacent when a tensor holds both),
# min-degree / min-fill elimination order on Python-int bitsets, then every eliminated index
# contracts the tensors that hold it.  On lattice x depth circuit networks this finds the
# "contract the time direction, then sweep the lattice with a boundary" trees that a greedy
# pairwise search misses; its cost is 2^(elimination width) per step.
def _elimination_order(net: _Net, masks: Sequence[int], output: int, rule: str = "min_fill",
                       seed: int = 0) -> List[int]:  # fmt: skip
    nsym = len(net.names)
    adj = [0] * nsym
    for m in masks:
        for i in _bits(m):
            adj[i] |= m
    for i in range(nsym):
        adj[i] &= ~(1 << i)
    alive = 0
    for m in masks:
        alive |= m
    elim = alive & ~output
    order: List[int] = []
    import random

    rnd = random.Random(seed)
    while elim:
        best = None
        cands = list(_bits(elim))
        # min degree first; among the lowest few, least fill
        degs = sorted((_popcount(adj[v] & alive), v) for v in cands)
        lo = degs[0][0]
        short = [v for d, v in degs if d <= lo + (1 if rule == "min_fill" else 0)][:24]
        for v in short:
            nb = adj[v] & alive
            if rule == "min_fill":
                fill = 0
                for u in _bits(nb):
                    fill += _popcount(nb & ~adj[u]) - 1
                key = (fill, _popcount(nb), rnd.random() if seed else 0, v)
            else:
                key = (_popcount(nb), rnd.random() if seed else 0, v)
            if best is None or key < best[0]:
                best = (key, v)
        v = best[1]
        nb = adj[v] & alive
        for u in _bits(nb):
            adj[u] |= nb
            adj[u] &= ~((1 << u) | (1 << v))
        alive &= ~(1 << v)
        elim &= ~(1 << v)
        order.append(v)
    return order


def _order_to_ssa(net: _Net, order: Sequence[int], terms: Dict[int, int], output: int, nxt: int) -> List[Tuple[int, int]]:
    ssa: List[Tuple[int, int]] = []
    occ = [0] * len(net.names)
    where: List[set] = [set() for _ in range(len(net.names))]
    for t, m in terms.items():
        for i in _bits(m):
            occ[i] += 1
            where[i].add(t)

    def size_after(a: int, b: int) -> int:
        ma, mb = terms[a], terms[b]
        union, both = ma | mb, ma & mb
        n = _popcount(union & output)
        for i in _bits(union & ~output):
            if occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                n += 1
        return n

    def merge(a: int, b: int) -> int:
        nonlocal nxt
        ma, mb = terms[a], terms[b]
        union, both = ma | mb, ma & mb
        keep = union & output
        for i in _bits(union & ~output):
            if occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                keep |= 1 << i
        for t, m in ((a, ma), (b, mb)):
            for i in _bits(m):
                occ[i] -= 1
                where[i].discard(t)
            del terms[t]
        terms[nxt] = keep
        for i in _bits(keep):
            occ[i] += 1
            where[i].add(nxt)
        ssa.append((a, b))
        nxt += 1
        return nxt - 1

    for v in order:
        live = sorted(where[v])
        while len(live) > 1:
            best = None
            for x in range(len(live)):
                for y in range(x + 1, len(live)):
                    k = (size_after(live[x], live[y]), live[x], live[y])
                    if best is None or k < best[0]:
                        best = (k, x, y)
            _, x, y = best
            new = merge(live[x], live[y])
            live = [t for k, t in enumerate(live) if k not in (x, y)]
            if (terms[new] >> v) & 1:
                live.append(new)
    # whatever is left shares only output indices (or nothing): smallest first
    rest = sorted(terms, key=lambda t: (_popcount(terms[t]), t))
    while len(rest) > 1:
        merge(rest[0], rest[1])
        rest = sorted(terms, key=lambda t: (_popcount(terms[t]), t))
    return ssa


def _wire_graph(net: _Net, masks: Sequence[int], groups: Dict[str, Any]):
    import networkx as nx

    gid = [groups.get(nm, ("_", nm)) for nm in net.names]
    wg = nx.Graph()
    wg.add_nodes_from(sorted(set(gid), key=str))
    for m in masks:
        gs = sorted({gid[i] for i in _bits(m)}, key=str)
        for a in range(len(gs)):
            for b in range(a + 1, len(gs)):
                w = wg[gs[a]][gs[b]]["weight"] + 1 if wg.has_edge(gs[a], gs[b]) else 1
                wg.add_edge(gs[a], gs[b], weight=w)
    return gid, wg


def _wire_orders(wg: Any, nangles: int = 16) -> List[List[Any]]:
    """Candidate linear arrangements of the qubit wires: reverse Cuthill-McKee, and sweeps along
    directions of the 2D spectral embedding (for a lattice these are the row / column / diagonal
    sweeps; the cheapest is picked by the caller, which costs the resulting tree)."""
    import networkx as nx
    import numpy as np
    from networkx.utils import reverse_cuthill_mckee_ordering

    nodes = list(wg.nodes)
    out: List[List[Any]] = []
    rcm: List[Any] = []
    for comp in nx.connected_components(wg):
        rcm += list(reverse_cuthill_mckee_ordering(wg.subgraph(comp)))
    out.append(rcm)
    if len(nodes) >= 4 and nx.is_connected(wg):
        lap = nx.laplacian_matrix(wg, nodelist=nodes, weight=None).toarray().astype(float)
        vals, vecs = np.linalg.eigh(lap)
        v1, v2 = vecs[:, 1], vecs[:, 2]
        for a in range(nangles):
            th = np.pi * a / nangles
            f = np.cos(th) * v1 + np.sin(th) * v2
            g = -np.sin(th) * v1 + np.cos(th) * v2
            # quantise the sweep coordinate into ~sqrt(n) levels so that a whole row / column ties and
            # is ordered by the orthogonal coordinate (a slightly tilted sweep direction would
            # interleave neighbouring rows and double the cut)
            fn = (f - f.min()) / (f.max() - f.min() + 1e-30)
            for levels in (int(round(len(nodes) ** 0.5)) - 1, 2 * int(round(len(nodes) ** 0.5)) - 2):
                key = np.round(fn * max(1, levels))
                order = sorted(range(len(nodes)), key=lambda k: (key[k], g[k]))
                cand = [nodes[k] for k in order]
                if cand not in out:
                    out.append(cand)
    return out


def _wire_sweep_order(net: _Net, masks: Sequence[int], output: int, gid: Sequence[Any], order_w: Sequence[Any]) -> List[int]:
    """Elimination order for circuit networks: indices are grouped by the qubit wire they belong to;
    wires are eliminated one after the other in `order_w`, so the boundary tensor only carries the
    bonds between finished and unfinished wires (a PEPS boundary sweep)."""
    rank = {w: k for k, w in enumerate(order_w)}
    alive = 0
    for m in masks:
        alive |= m
    idx = [i for i in _bits(alive & ~output)]
    idx.sort(key=lambda i: (rank[gid[i]], i))
    return idx


def search_elimination(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
                       target_size: Optional[int] = None, rules: Sequence[Tuple[str, int]] = (("wires", 0), ("min_fill", 0), ("min_degree", 0), ("min_fill", 1)),
                       max_slices_log2: int = 40, groups: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:  # fmt: skip
    """Simplify + variable elimination + greedy slicing; returns a `tree_data` dict.
    `groups` (symbol -> qubit wire id, cons.wire_groups) enables the wire-sweep order."""
    inputs = [tuple(t) for t in inputs]
    output = tuple(output)
    net = _Net(inputs, output, size_dict)
    best: Optional[Tuple[float, List[Tuple[int, int]], List[str]]] = None
    base_terms = {i: m for i, m in enumerate(net.inputs)}
    ssa0, nxt0 = _simplify_ssa(net, base_terms, net.output, len(inputs))
    candidates: List[List[int]] = []
    for rule, seed in rules:
        if rule == "wires":
            if groups:
                gid, wg = _wire_graph(net, list(base_terms.values()), groups)
                for order_w in _wire_orders(wg):
                    candidates.append(_wire_sweep_order(net, list(base_terms.values()), net.output, gid, order_w))
        else:
            candidates.append(_elimination_order(net, list(base_terms.values()), net.output, rule, seed))
    # cost every candidate tree without slicing first (cheap), slice only the few best
    scored = []
    for order in candidates:
        terms = dict(base_terms)
        ssa1 = _order_to_ssa(net, order, terms, net.output, nxt0)
        path = ssa_to_linear(ssa0 + ssa1, len(inputs))
        st = path_stats(inputs, output, size_dict, path)
        scored.append((st["flops"], st["size"], path))
    scored.sort(key=lambda x: (x[0], x[1]))
    for _, _, path in scored[:3]:
        sliced: List[str] = []
        st = path_stats(inputs, output, size_dict, path)
        if target_size is not None:
            while st["size"] > target_size and len(sliced) < max_slices_log2:
                s = _pick_slice_index(inputs, output, size_dict, path, sliced, target_size)
                if s is None:
                    break
                sliced.append(s)
                st = path_stats(inputs, output, size_dict, path, sliced)
        total = st["flops"] * st["nslices"]
        if target_size is not None and st["size"] > target_size:
            total *= 1e30
        if best is None or total < best[0]:
            best = (total, path, list(sliced))
    assert best is not None
    _, path, sliced = best
    return {
        "inputs": tuple(inputs),
        "output": output,
        "size_dict": dict(size_dict),
        "path": [tuple(p) for p in path],
        "sliced_inds": {s: size_dict[s] for s in sliced},
        "planner": "tensorcircuit_ng_b200.planner.search_elimination (ours; not cotengra)",
    }


# ---------------------------------------------------------------------------------------------
# Site-block sweep planner (round 2).  The wire-sweep elimination above removes ONE index at a time, so its
# trees are made of skinny absorptions (a big tensor times a 2..64-element one) and its boundary carries every
# half-finished wire: 2^57 slices on the 7x7 depth-20 amplitude.  Here the network is coarsened first —
# every qubit wire's own tensors (inputs, one-qubit gates, closing one-hots) are contracted along the time
# direction into ONE site tensor over the hyper-indices its couplers touch — and the sites are then absorbed
# into a boundary tensor one at a time, couplers to already absorbed sites folded into the site tensor just
# before (a PEPS boundary contraction).  Every big step is a boundary x site contraction over all the bonds
# between them: GEMM shaped (K = 2^#bonds), which is what the tcgen05 kernel wants.  Site orders tried: RCM /
# spectral sweeps of the wire graph and greedy boundary growth from low-degree corners; the cheapest sliced
# tree wins.  This is the role cotengra's partition-based trees + slicing play for the reference
# (tensorcircuit/experimental.py:930-954); cotengra itself is not installed here.
class _Ssa:
    def __init__(self, net: _Net, terms: Dict[int, int], nxt: int) -> None:
        self.net, self.terms, self.nxt, self.out = net, terms, nxt, net.output
        self.ssa: List[Tuple[int, int]] = []
        self.occ = [0] * len(net.names)
        for m in terms.values():
            for i in _bits(m):
                self.occ[i] += 1

    def result(self, a: int, b: int) -> int:
        ma, mb = self.terms[a], self.terms[b]
        union, both = ma | mb, ma & mb
        keep = union & self.out
        for i in _bits(union & ~self.out):
            if self.occ[i] - (2 if (both >> i) & 1 else 1) > 0:
                keep |= 1 << i
        return keep

    def merge(self, a: int, b: int) -> int:
        keep = self.result(a, b)
        for t in (a, b):
            for i in _bits(self.terms[t]):
                self.occ[i] -= 1
            del self.terms[t]
        for i in _bits(keep):
            self.occ[i] += 1
        self.terms[self.nxt] = keep
        self.ssa.append((a, b))
        self.nxt += 1
        return self.nxt - 1

    def merge_all(self, ids: List[int]) -> int:
        """Greedy by result size inside a small group."""
        live = list(ids)
        while len(live) > 1:
            best = None
            for x in range(len(live)):
                for y in range(x + 1, len(live)):
                    if not (self.terms[live[x]] & self.terms[live[y]]) and len(live) > 2:
                        continue
                    k = (_popcount(self.result(live[x], live[y])), live[x], live[y])
                    if best is None or k < best[0]:
                        best = (k, x, y)
            if best is None:  # nothing shares an index: outer product of the two smallest
                live.sort(key=lambda t: (_popcount(self.terms[t]), t))
                best = ((0, 0, 0), 0, 1)
            _, x, y = best
            new = self.merge(live[x], live[y])
            live = [t for k, t in enumerate(live) if k not in (x, y)] + [new]
        return live[0]


def _site_tree(net: _Net, gid: Sequence[Any], order_w: Sequence[Any], n_inputs: int) -> List[Tuple[int, int]]:
    b = _Ssa(net, {i: m for i, m in enumerate(net.inputs)}, n_inputs)
    wires_of = {t: frozenset(gid[i] for i in _bits(m)) for t, m in b.terms.items()}
    own: Dict[Any, List[int]] = {}
    couplers: List[int] = []
    scalars: List[int] = []
    for t, ws in wires_of.items():
        if len(ws) == 1:
            own.setdefault(next(iter(ws)), []).append(t)
        elif len(ws) == 0:
            scalars.append(t)
        else:
            couplers.append(t)
    site: Dict[Any, int] = {}
    for w in order_w:
        if w in own:
            site[w] = b.merge_all(sorted(own[w]))
    done: set = set()
    boundary: Optional[int] = None
    pending = {c: wires_of[c] for c in couplers}
    for w in order_w:
        done.add(w)
        t = site.get(w)
        ready = sorted(c for c, ws in pending.items() if w in ws and ws <= done)
        for c in ready:
            del pending[c]
            t = c if t is None else b.merge(t, c)
        if t is None:
            continue
        boundary = t if boundary is None else b.merge(boundary, t)
    rest = [x for x in list(pending) + scalars]
    for x in rest:
        boundary = x if boundary is None else b.merge(boundary, x)
    return b.ssa


def _greedy_site_orders(wg: Any, starts: int = 4) -> List[List[Any]]:
    """Grow the absorbed set one site at a time, always taking the site that leaves the smallest cut."""
    nodes = sorted(wg.nodes, key=str)
    if not nodes:
        return []
    deg = {v: sum(d.get("weight", 1) for _, _, d in wg.edges(v, data=True)) for v in nodes}
    firsts = sorted(nodes, key=lambda v: (deg[v], str(v)))[:starts]
    out = []
    for f in firsts:
        inside = {f}
        order = [f]
        cut = {v: 0 for v in nodes}  # weight of edges from v into `inside`
        for _, u, d in wg.edges(f, data=True):
            cut[u] += d.get("weight", 1)
        while len(order) < len(nodes):
            best = None
            for v in nodes:
                if v in inside:
                    continue
                delta = deg[v] - 2 * cut[v]  # change of the cut when v joins
                key = (0 if cut[v] > 0 else 1, delta, -cut[v], str(v))
                if best is None or key < best[0]:
                    best = (key, v)
            v = best[1]
            inside.add(v)
            order.append(v)
            for _, u, d in wg.edges(v, data=True):
                if u not in inside:
                    cut[u] += d.get("weight", 1)
        out.append(order)
    return out


def search_sites(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
                 groups: Dict[str, Any], target_size: Optional[int] = None, max_slices_log2: int = 48,
                 keep: int = 3) -> Dict[str, Any]:  # fmt: skip
    """Site-block sweep + greedy slicing; returns a `tree_data` dict (tensorcircuit/experimental.py:947-953)."""
    inputs = [tuple(t) for t in inputs]
    output = tuple(output)
    net = _Net(inputs, output, size_dict)
    gid, wg = _wire_graph(net, net.inputs, groups)
    orders = _greedy_site_orders(wg) + _wire_orders(wg)
    scored = []
    seen = set()
    for order_w in orders:
        key = tuple(str(w) for w in order_w)
        if key in seen:
            continue
        seen.add(key)
        path = ssa_to_linear(_site_tree(net, gid, order_w, len(inputs)), len(inputs))
        st = path_stats(inputs, output, size_dict, path)
        scored.append((st["flops"], st["size"], path))
    scored.sort(key=lambda x: (x[0], x[1]))
    best: Optional[Tuple[float, List[Tuple[int, int]], List[str]]] = None
    for _, _, path in scored[:keep]:
        sliced: List[str] = []
        st = path_stats(inputs, output, size_dict, path)
        if target_size is not None:
            while st["size"] > target_size and len(sliced) < max_slices_log2:
                s = _pick_slice_index(inputs, output, size_dict, path, sliced, target_size)
                if s is None:
                    break
                sliced.append(s)
                st = path_stats(inputs, output, size_dict, path, sliced)
        total = st["flops"] * st["nslices"]
        if target_size is not None and st["size"] > target_size:
            total *= 1e30
        if best is None or total < best[0]:
            best = (total, path, list(sliced))
    assert best is not None
    _, path, sliced = best
    return {
        "inputs": tuple(inputs),
        "output": output,
        "size_dict": dict(size_dict),
        "path": [tuple(p) for p in path],
        "sliced_inds": {s: size_dict[s] for s in sliced},
        "planner": "tensorcircuit_ng_b200.planner.search_sites (ours; not cotengra)",
    }
