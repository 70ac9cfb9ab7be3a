"""
ORACLE (test infrastructure, never imported by the product package).

Restatement of the two `opt_einsum==3.4.0` path finders the reference calls
(/root/reference/tensorcircuit/cons.py:1020 `opt_einsum.paths.optimal`,
:1246 `getattr(opt_einsum.paths, "greedy")`).  opt_einsum is a third-party
dependency (pinned in requirements/requirements-2411.txt) that is absent from
this image; the published algorithm is restated from memory
[UPSTREAM-UNVERIFIED].  PARITY UNPINNED for bit-exact tie-breaking: no reference
test fixes a concrete path; numerics are path independent up to fp32 rounding.
The one plan-level anchor is the published greedy cost triple for
`example_block` n=10, 4 layers (docs/source/tutorials/contractors.ipynb):
log10 FLOPs 5.132, log2 SIZE 11, log2 WRITE 13.083 — checked in
tests/test_oracle_golden.py.

Path convention: opt_einsum *linear* paths — pop both operands (higher index
first), append the result (consumed at cons.py:937-950).
"""

from __future__ import annotations

import heapq
import itertools
from collections import defaultdict
from typing import Any, Dict, FrozenSet, List, Optional, Sequence, Tuple


def compute_size_by_dict(indices, idx_dict: Dict[str, int]) -> int:
    ret = 1
    for i in indices:
        ret *= idx_dict[i]
    return ret


def ssa_to_linear(ssa_path: Sequence[Tuple[int, ...]]) -> List[Tuple[int, ...]]:
    n = 1 + max(map(max, ssa_path))
    ids = list(range(n))
    path = []
    for ssa_ids in ssa_path:
        path.append(tuple(int(ids[s]) for s in ssa_ids))
        for s in ssa_ids:
            for k in range(s, n):
                ids[k] -= 1
    return path


def _get_candidate(output, sizes, remaining, footprints, dim_ref_counts, k1, k2):
    either = k1 | k2
    two = k1 & k2
    one = either - two
    k12 = (either & output) | (two & dim_ref_counts[3]) | (one & dim_ref_counts[2])
    cost = compute_size_by_dict(k12, sizes) - footprints[k1] - footprints[k2]  # "memory-removed"
    id1 = remaining[k1]
    id2 = remaining[k2]
    if id1 > id2:
        k1, id1, k2, id2 = k2, id2, k1, id1
    return (cost, id2, id1), k1, k2, k12


def _push_candidate(output, sizes, remaining, footprints, dim_ref_counts, k1, k2s, queue):
    cands = [_get_candidate(output, sizes, remaining, footprints, dim_ref_counts, k1, k2) for k2 in k2s]
    best = min(cands, key=lambda c: c[0])
    heapq.heappush(queue, _HeapItem(best))


class _HeapItem:
    """Order candidates by their (cost, id2, id1) key only (frozensets are not totally ordered)."""

    __slots__ = ("c",)

    def __init__(self, c):
        self.c = c

    def __lt__(self, other):
        return self.c[0] < other.c[0]


def _update_ref_counts(dim_to_keys, dim_ref_counts, dims):
    for dim in dims:
        count = len(dim_to_keys[dim])
        if count <= 1:
            dim_ref_counts[2].discard(dim)
            dim_ref_counts[3].discard(dim)
        elif count == 2:
            dim_ref_counts[2].add(dim)
            dim_ref_counts[3].discard(dim)
        else:
            dim_ref_counts[2].add(dim)
            dim_ref_counts[3].add(dim)


def ssa_greedy_optimize(inputs, output, sizes) -> List[Tuple[int, ...]]:
    if len(inputs) == 1:
        return [(0,)]
    fs_inputs = [frozenset(x) for x in inputs]
    output = frozenset(output) | frozenset.intersection(*fs_inputs)

    remaining: Dict[FrozenSet[str], int] = {}
    ssa_ids = itertools.count(len(fs_inputs))
    ssa_path: List[Tuple[int, ...]] = []
    for ssa_id, key in enumerate(fs_inputs):
        if key in remaining:
            ssa_path.append((remaining[key], ssa_id))
            remaining[key] = next(ssa_ids)
        else:
            remaining[key] = ssa_id

    dim_to_keys = defaultdict(set)
    for key in remaining:
        for dim in key - output:
            dim_to_keys[dim].add(key)

    dim_ref_counts = {
        count: set(dim for dim, keys in dim_to_keys.items() if len(keys) >= count) - output
        for count in [2, 3]
    }
    footprints = {key: compute_size_by_dict(key, sizes) for key in remaining}

    queue: List[_HeapItem] = []
    for dim, keys in dim_to_keys.items():
        keys = sorted(keys, key=remaining.__getitem__)
        for i, k1 in enumerate(keys[:-1]):
            k2s = keys[1 + i :]
            _push_candidate(output, sizes, remaining, footprints, dim_ref_counts, k1, k2s, queue)

    while queue:
        cost, k1, k2, k12 = heapq.heappop(queue).c
        if k1 not in remaining or k2 not in remaining:
            continue
        ssa_id1 = remaining.pop(k1)
        ssa_id2 = remaining.pop(k2)
        for dim in k1 - output:
            dim_to_keys[dim].remove(k1)
        for dim in k2 - output:
            dim_to_keys[dim].remove(k2)
        ssa_path.append((ssa_id1, ssa_id2))
        if k12 in remaining:
            ssa_path.append((remaining[k12], next(ssa_ids)))
        else:
            for dim in k12 - output:
                dim_to_keys[dim].add(k12)
        remaining[k12] = next(ssa_ids)
        _update_ref_counts(dim_to_keys, dim_ref_counts, k1 | k2 - output)
        footprints[k12] = compute_size_by_dict(k12, sizes)

        k1 = k12
        k2s = set(k2 for dim in k1 - output for k2 in dim_to_keys[dim])
        k2s.discard(k1)
        if k2s:
            k2s_sorted = sorted(k2s, key=remaining.__getitem__)
            _push_candidate(output, sizes, remaining, footprints, dim_ref_counts, k1, k2s_sorted, queue)

    # outer products of whatever is left, smallest first
    rest = [(compute_size_by_dict(key & output, sizes), ssa_id, key) for key, ssa_id in remaining.items()]
    heapq.heapify(rest)
    _, ssa_id1, k1 = heapq.heappop(rest)
    while rest:
        _, ssa_id2, k2 = heapq.heappop(rest)
        ssa_path.append((min(ssa_id1, ssa_id2), max(ssa_id1, ssa_id2)))
        k12 = (k1 | k2) & output
        cost = compute_size_by_dict(k12, sizes)
        ssa_id12 = next(ssa_ids)
        _, ssa_id1, k1 = heapq.heappushpop(rest, (cost, ssa_id12, k12))
    return ssa_path


def greedy(inputs, output, size_dict, memory_limit: Optional[int] = None) -> List[Tuple[int, ...]]:
    """opt_einsum.paths.greedy, default cost "memory-removed"."""
    ssa = ssa_greedy_optimize(inputs, output, size_dict)
    return ssa_to_linear(ssa)


def _flop_count(idx_contraction, inner: bool, num_terms: int, size_dict) -> int:
    overall = compute_size_by_dict(idx_contraction, size_dict)
    op_factor = max(1, num_terms - 1)
    if inner:
        op_factor += 1
    return overall * op_factor


def optimal(inputs, output, size_dict, memory_limit: Optional[int] = None) -> List[Tuple[int, ...]]:
    """opt_einsum.paths.optimal: exhaustive DFS over pair orders minimising flops
    (used by the reference when len(nodes) < 5, cons.py:1019-1030)."""
    inputs = tuple(frozenset(x) for x in inputs)
    output = frozenset(output)
    best = {"flops": float("inf"), "ssa_path": tuple((i,) for i in range(len(inputs)))}
    size_cache: Dict[Any, int] = {}
    result_cache: Dict[Any, Any] = {}

    def _iterate(path, remaining, inputs_, flops):
        if len(remaining) == 1:
            best["flops"] = flops
            best["ssa_path"] = path
            return
        for i, j in itertools.combinations(remaining, 2):
            if i > j:
                i, j = j, i
            key = (inputs_[i], inputs_[j])
            try:
                k12, flops12 = result_cache[key]
            except KeyError:
                either = inputs_[i] | inputs_[j]
                rest = [inputs_[r] for r in remaining if r not in (i, j)]
                keep = frozenset.union(output, *rest) if rest else output
                k12 = either & keep
                flops12 = _flop_count(either, bool(either - k12), 2, size_dict)
                result_cache[key] = (k12, flops12)
            # the kept set depends on what else remains; recompute exactly
            rest = [inputs_[r] for r in remaining if r not in (i, j)]
            keep = frozenset.union(output, *rest) if rest else output
            either = inputs_[i] | inputs_[j]
            k12 = either & keep
            flops12 = _flop_count(either, bool(either - k12), 2, size_dict)
            new_flops = flops + flops12
            if new_flops >= best["flops"]:
                continue
            _iterate(
                path + ((i, j),),
                remaining - {i, j} | {len(inputs_)},
                inputs_ + (k12,),
                new_flops,
            )

    _iterate((), set(range(len(inputs))), inputs, 0)
    if len(inputs) == 1:
        return [(0,)]
    return ssa_to_linear(best["ssa_path"])


def path_cost(inputs, output, size_dict, path) -> Dict[str, float]:
    """FLOPs / SIZE / WRITE of a linear path in opt_einsum/cotengra accounting:
    flops = sum over steps of prod(sizes of all indices involved) (scalar MACs),
    write = sum of output sizes (+ leaves are not counted), size = largest intermediate."""
    terms = [frozenset(x) for x in inputs]
    out = frozenset(output)
    flops = 0
    write = 0
    size = 0
    for step in path:
        if len(step) < 2:
            continue
        i, j = step
        a, b = terms[i], terms[j]
        rest = [t for k, t in enumerate(terms) if k not in (i, j)]
        keep = frozenset.union(out, *rest) if rest else out
        k12 = (a | b) & keep
        flops += compute_size_by_dict(a | b, size_dict)
        s = compute_size_by_dict(k12, size_dict)
        write += s
        size = max(size, s)
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
        terms.append(k12)
    return {"flops": flops, "write": write, "size": size}
