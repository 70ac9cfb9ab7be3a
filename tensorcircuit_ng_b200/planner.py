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
        tot, i = 0.0, 0
        while m:
            if m & 1:
                tot += self.log2[i]
            m >>= 1
            i += 1
        return tot


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
        i = 0
        mm = m
        while mm:
            if mm & 1:
                occ[i] += 1
                where[i].add(tid)
            mm >>= 1
            i += 1

    def result_of(a: int, b: int) -> int:
        ma, mb = terms[a], terms[b]
        both = ma & mb
        keep = (ma | mb) & output
        rest = (ma | mb) & ~output
        i = 0
        mm = rest
        while mm:
            if mm & 1:
                need = occ[i] - (1 if (ma >> i) & 1 else 0) - (1 if (mb >> i) & 1 else 0)
                if need > 0:
                    keep |= 1 << i
            mm >>= 1
            i += 1
        return keep

    def cost(a: int, b: int) -> Tuple[float, int, int]:
        r = result_of(a, b)
        c = 2.0 ** net.lsize(r) - alpha * (2.0 ** net.lsize(terms[a]) + 2.0 ** net.lsize(terms[b]))
        lo, hi = (a, b) if a < b else (b, a)
        return (c, hi, lo)

    heap: List[Tuple[Tuple[float, int, int], int, int]] = []

    def push_neighbours(t: int) -> None:
        nb = set()
        mm = terms[t]
        i = 0
        while mm:
            if mm & 1:
                nb.update(where[i])
            mm >>= 1
            i += 1
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
            mm = terms[t]
            i = 0
            while mm:
                if mm & 1:
                    occ[i] -= 1
                    where[i].discard(t)
                mm >>= 1
                i += 1
            del terms[t]
        new = nxt
        nxt += 1
        terms[new] = r
        mm = r
        i = 0
        while mm:
            if mm & 1:
                occ[i] += 1
                where[i].add(new)
            mm >>= 1
            i += 1
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
    write = sum of intermediate sizes, size = largest intermediate, all as plain numbers."""
    sl = set(sliced)
    terms = [[s for s in t if s not in sl] for t in inputs]
    out = [s for s in output if s not in sl]
    flops = 0.0
    write = 0.0
    size = 0.0
    per_step: List[Tuple[int, int, int, int]] = []
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        rest = set(out)
        for k, t in enumerate(terms):
            if k not in (i, j):
                rest.update(t)
        allidx = list(dict.fromkeys(a + b))
        keep = [s for s in allidx if s in rest]
        f = 1.0
        for s in allidx:
            f *= size_dict[s]
        w = 1.0
        for s in keep:
            w *= size_dict[s]
        flops += f
        write += w
        size = max(size, w)
        per_step.append((len(a), len(b), len(keep), len(allidx)))
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
        terms.append(keep)
    nslices = 1.0
    for s in sliced:
        nslices *= size_dict[s]
    return {"flops": flops, "write": write, "size": size, "nslices": nslices, "steps": per_step}  # type: ignore[dict-item]


def greedy(inputs: Sequence[Sequence[str]], output: Sequence[str], size_dict: Dict[str, int],
           memory_limit: Optional[int] = None) -> List[Tuple[int, int]]:  # fmt: skip
    """Optimizer callable with the reference's Level-1 plug signature."""
    if len(inputs) == 1:
        return []
    net = _Net(inputs, output, size_dict)
    return ssa_to_linear(_ssa_greedy(net), len(inputs))


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
    sl = set(sliced)
    terms = [[s for s in t if s not in sl] for t in inputs]
    out = [s for s in output if s not in sl]
    score: Dict[str, float] = {}
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        rest = set(out)
        for k, t in enumerate(terms):
            if k not in (i, j):
                rest.update(t)
        allidx = list(dict.fromkeys(a + b))
        keep = [s for s in allidx if s in rest]
        w = 1.0
        for s in keep:
            w *= size_dict[s]
        f = 1.0
        for s in allidx:
            f *= size_dict[s]
        if w > target_size:
            for s in keep:
                if s not in out:
                    score[s] = score.get(s, 0.0) + w * 1e6 + f
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
        terms.append(keep)
    if not score:
        return None
    return max(sorted(score), key=lambda s: score[s])


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
