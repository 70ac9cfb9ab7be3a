"""
ORACLE (test infrastructure / CPU baseline only) — the reference's pairwise tree execution on numpy.

`tensorcircuit/cons.py:937-953` (`_base`): for every (i, j) of the path, contract the two nodes over their
shared edges (`tn.contract_between` -> one `tensordot`), append the result, drop the operands; sliced plans
(`tensorcircuit/experimental.py:999-1009`) first fix the sliced indices of every leaf.  Restated on plain
arrays and index tuples (the `tree_data` schema of `:947-953`).  Hyper-indices (an index carried by more than
two tensors, what cotengra executes through einsum) stay until no other tensor and no output needs them.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def contract_tree_numpy(arrays: Sequence[np.ndarray], inputs: Sequence[Sequence[str]], output: Sequence[str],
                        path: Sequence[Tuple[int, int]], fixed: Optional[Dict[str, int]] = None) -> np.ndarray:  # fmt: skip
    terms: List[Tuple[np.ndarray, Tuple[str, ...]]] = []
    for a, ix in zip(arrays, inputs):
        a = np.asarray(a)
        ix = tuple(ix)
        if fixed:
            sel = tuple(fixed[s] if s in fixed else slice(None) for s in ix)
            a = a[sel]
            ix = tuple(s for s in ix if s not in fixed)
        terms.append((a, ix))
    out = tuple(s for s in output if not (fixed and s in fixed))
    for i, j in path:
        i, j = (i, j) if i < j else (j, i)
        b, ixb = terms.pop(j)
        a, ixa = terms.pop(i)
        needed = set(out)
        for _, ix in terms:
            needed.update(ix)
        keep = [s for s in ixa if s in needed] + [s for s in ixb if s in needed and s not in ixa]
        sym = {s: k for k, s in enumerate(dict.fromkeys(ixa + ixb))}
        r = np.einsum(a, [sym[s] for s in ixa], b, [sym[s] for s in ixb], [sym[s] for s in keep], optimize=True)
        terms.append((r, tuple(keep)))
    r, ix = terms[0]
    for extra, ixe in terms[1:]:  # disconnected pieces (scalars)
        r = np.multiply.outer(r, extra)
        ix = ix + ixe
    if ix != out:
        r = np.transpose(r, [ix.index(s) for s in out])
    return r
