"""
Graph-level simplifications applied before the contractor sees a network
(/root/reference/tensorcircuit/simplify.py:83-85 `_multi_remove`, :198-296 light-cone
cancellation used by `expectation(enable_lightcone=True)`, tensorcircuit/circuit.py:897-901).
Pure graph surgery on the duck-typed node surface; no arithmetic.
"""

from __future__ import annotations

from typing import Any, List, Tuple


def _multi_remove(elems: List[Any], indices: List[int]) -> List[Any]:
    drop = set(indices)
    return [e for i, e in enumerate(elems) if i not in drop]


def _light_cone_cancel(nodes: List[Any]) -> Tuple[List[Any], bool]:
    """One backward scan: a gate U whose every output leg meets the same leg of its own
    conjugate copy U^dagger is bypassed (U^dagger U = 1)."""
    changed = False
    removed = set()
    for n in reversed(nodes):
        if id(n) in removed or getattr(n, "is_dagger", None) is True:
            continue
        rank = len(n.shape)
        if rank % 2:
            continue
        half = rank // 2
        partner = None
        ok = True
        for leg in range(half):
            e = n[leg]
            if e.is_dangling():
                ok = False
                break
            other = e.node2 if e.node1 is n else e.node1
            if (
                getattr(other, "is_dagger", None) is not True
                or getattr(other, "id", None) != getattr(n, "id", -1)
                or e.axis1 != e.axis2
                or (partner is not None and partner is not other)
            ):
                ok = False
                break
            partner = other
        if not ok or partner is None or id(partner) in removed:
            continue
        for leg in range(half, rank):
            e_n, e_m = n[leg], partner[leg]
            m_n, i_n = (e_n.node2, e_n.axis2) if e_n.node1 is n else (e_n.node1, e_n.axis1)
            m_m, i_m = (e_m.node2, e_m.axis2) if e_m.node1 is partner else (e_m.node1, e_m.axis1)
            e_n.disconnect()
            e_m.disconnect()
            m_n[i_n] ^ m_m[i_m]
        removed.add(id(n))
        removed.add(id(partner))
        changed = True
    if changed:
        return [x for x in nodes if id(x) not in removed], True
    return nodes, False


def _full_light_cone_cancel(nodes: List[Any]) -> List[Any]:
    if not nodes or any(getattr(n, "is_dagger", None) is None for n in nodes):
        return nodes
    nodes, changed = _light_cone_cancel(nodes)
    while changed:
        nodes, changed = _light_cone_cancel(nodes)
    return nodes
