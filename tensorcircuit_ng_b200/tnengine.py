"""
Tensor-network executor: pairwise contractions and whole contraction trees through
`tcb_tn_contract` (include/tcb200.h).

Replaces `tn.contract_between -> backend.tensordot` (tensorcircuit/cons.py:948) and
cotengra's `contract_core` (tensorcircuit/experimental.py:1008).  Because every mode has
extent 2, an operand's layout is just "which flat-index bit each mode occupies": permuted
views (power-of-two strides), lazily conjugated tensors (torch conj bit) and sliced leaves
(a fixed bit => element offset, K7) are all consumed *in place* — nothing is transposed,
conjugated or sliced into a temporary.

The torch autograd vjp (`_Contract.backward`) issues the same kernel for the two reverse
contractions  dA = dC . B^H,  dB = A^H . dC   (north-star kernel (3)).
"""

from __future__ import annotations

import math
from typing import Any, Dict, Hashable, List, Optional, Sequence, Tuple

import torch

from . import _lib

Mode = Hashable


def _bitpos_of(t: torch.Tensor) -> Optional[List[int]]:
    """Flat-index bit position of every axis of a dim-2 tensor view, or None if not expressible."""
    pos = []
    for size, stride in zip(t.shape, t.stride()):
        if size != 2:
            raise _lib.EngineError(
                f"tensor of shape {tuple(t.shape)}: only qubit (extent-2) modes are supported by the B200 engine"
            )
        if stride <= 0 or stride & (stride - 1):
            return None
        pos.append(stride.bit_length() - 1)
    if len(set(pos)) != len(pos):
        return None
    return pos


def _physical(t: torch.Tensor) -> Tuple[torch.Tensor, List[int], int, int]:
    """(storage-backed tensor, bit position per axis, conj flag, element offset of the view)."""
    if t.dtype != torch.complex64:
        t = t.to(torch.complex64)
    conj = 0
    if t.is_conj():
        t = t.conj()  # flips the lazy bit back: a plain view of the un-conjugated memory
        conj = 1
    pos = _bitpos_of(t)
    if pos is None:
        t = t.contiguous()
        pos = _bitpos_of(t)
        assert pos is not None
    return t, pos, conj, 0


def row_side_first(modes_a: Sequence[Mode], modes_b: Sequence[Mode], modes_out: Sequence[Mode]) -> bool:
    """True when (a, b) may stay in this order: the larger free side is the row side (A) — streaming and the
    128-row tensor-core tiles run along it."""
    out_set = set(modes_out)
    free_a = sum(1 for m in modes_a if m in out_set and m not in modes_b)
    free_b = sum(1 for m in modes_b if m in out_set and m not in modes_a)
    return free_b <= free_a


def make_desc(pa: Dict[Mode, int], pb: Dict[Mode, int], modes_a: Sequence[Mode], modes_b: Sequence[Mode],
              modes_out: Sequence[Mode], fixed: Any, conj_a: bool, conj_b: bool) -> "_lib.ContractDesc":  # fmt: skip
    """The `tcb_contract_desc` of one pairwise step from the bit position of every mode in the two operands
    (`pa`, `pb`); the result is contiguous in `modes_out` order; modes in `fixed` are sliced (skipped)."""
    n_out = len(modes_out)
    pc = {m: n_out - 1 - i for i, m in enumerate(modes_out)}
    d = _lib.ContractDesc()
    lists: Dict[str, List[int]] = {k: [] for k in ("batch_a", "batch_b", "batch_c", "m_a", "m_c", "n_b", "n_c", "k_a", "k_b")}
    for m in modes_out:
        ina, inb = m in pa, m in pb
        if m in fixed:
            raise _lib.EngineError(f"sliced mode {m!r} cannot be an output mode")
        if ina and inb:
            lists["batch_a"].append(pa[m]); lists["batch_b"].append(pb[m]); lists["batch_c"].append(pc[m])  # noqa: E702
        elif ina:
            lists["m_a"].append(pa[m]); lists["m_c"].append(pc[m])  # noqa: E702
        elif inb:
            lists["n_b"].append(pb[m]); lists["n_c"].append(pc[m])  # noqa: E702
        else:
            raise _lib.EngineError(f"output mode {m!r} is in neither operand")
    for m in modes_a:
        if m in pc or m in fixed:
            continue
        if m in pb:
            lists["k_a"].append(pa[m]); lists["k_b"].append(pb[m])  # noqa: E702
        else:
            raise _lib.EngineError(f"mode {m!r} appears only in the left operand and is not an output")
    for m in modes_b:
        if m not in pc and m not in fixed and m not in pa:
            raise _lib.EngineError(f"mode {m!r} appears only in the right operand and is not an output")
    # Inside a class the logical bit order is free (it only has to agree between the operands):
    # K bit i <-> i-th lowest position in A, N bit i <-> i-th lowest position in C, M likewise, so
    # that the kernels can use vector loads / stores when a class occupies an operand's lowest bits.
    def _sort(keys: Sequence[str], by: str) -> None:
        order = sorted(range(len(lists[by])), key=lambda i: lists[by][i])
        for kname in keys:
            lists[kname] = [lists[kname][i] for i in order]

    _sort(("k_a", "k_b"), "k_a")
    _sort(("n_b", "n_c"), "n_c")
    _sort(("m_a", "m_c"), "m_c")
    _sort(("batch_a", "batch_b", "batch_c"), "batch_c")
    for k, v in lists.items():
        if len(v) > 32:
            raise _lib.EngineError("too many modes in one class (max 32)")
        arr = getattr(d, k)
        for i, x in enumerate(v):
            arr[i] = x
    d.n_batch, d.n_m, d.n_n, d.n_k = len(lists["batch_c"]), len(lists["m_c"]), len(lists["n_c"]), len(lists["k_a"])
    d.conj_a = int(bool(conj_a))
    d.conj_b = int(bool(conj_b))
    return d


def contract_raw(a: torch.Tensor, modes_a: Sequence[Mode], b: torch.Tensor, modes_b: Sequence[Mode],
                 modes_out: Sequence[Mode], conj_a: bool = False, conj_b: bool = False,
                 fixed: Optional[Dict[Mode, int]] = None, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:  # fmt: skip
    """out[modes_out] (+)= sum over shared non-output modes of a[modes_a] * b[modes_b].

    `fixed` pins modes to a value (slicing, K7): a pinned mode is neither summed nor output.
    No autograd; see `contract` for the differentiable entry point.
    """
    fixed = fixed or {}
    # the larger free side becomes the row side (A): streaming / 128-row tiles run along it
    if not row_side_first(modes_a, modes_b, modes_out):
        a, b, modes_a, modes_b, conj_a, conj_b = b, a, modes_b, modes_a, conj_b, conj_a
    a, pos_a, cja, _ = _physical(a)
    b, pos_b, cjb, _ = _physical(b)
    _lib.require_cuda(a, "left operand")
    _lib.require_cuda(b, "right operand")
    pa = dict(zip(modes_a, pos_a))
    pb = dict(zip(modes_b, pos_b))
    if len(pa) != len(modes_a) or len(pb) != len(modes_b):
        raise _lib.EngineError("repeated mode inside one operand (trace) is not a pairwise contraction")
    n_out = len(modes_out)
    a_off = sum(fixed[m] << p for m, p in pa.items() if m in fixed)
    b_off = sum(fixed[m] << p for m, p in pb.items() if m in fixed)
    d = make_desc(pa, pb, modes_a, modes_b, modes_out, fixed, bool(conj_a) ^ bool(cja), bool(conj_b) ^ bool(cjb))
    if out is None:
        out = torch.empty([2] * n_out, dtype=torch.complex64, device=a.device)
        accumulate = False
    import ctypes

    _lib.call("tcb_tn_contract", a.data_ptr(), a_off, b.data_ptr(), b_off, out.data_ptr(), ctypes.byref(d),
              int(accumulate), _lib.stream_ptr())  # fmt: skip
    return out


class _Contract(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, a: torch.Tensor, b: torch.Tensor, modes_a: Any, modes_b: Any, modes_out: Any,
                fixed: Any) -> torch.Tensor:  # fmt: skip
        ctx.save_for_backward(a, b)
        ctx.meta = (tuple(modes_a), tuple(modes_b), tuple(modes_out), dict(fixed or {}))
        return contract_raw(a, modes_a, b, modes_b, modes_out, fixed=fixed)

    @staticmethod
    def backward(ctx: Any, gc: torch.Tensor):  # type: ignore[override]
        a, b = ctx.saved_tensors
        modes_a, modes_b, modes_out, fixed = ctx.meta
        if fixed:
            raise _lib.EngineError("autograd through sliced leaves is handled by the tree executor")
        ga = gb = None
        gc = gc.contiguous() if not gc.is_conj() else gc.resolve_conj().contiguous()
        if ctx.needs_input_grad[0]:
            ga = contract_raw(gc, modes_out, b, modes_b, modes_a, conj_b=True)
            ga = ga.reshape(a.shape)
        if ctx.needs_input_grad[1]:
            gb = contract_raw(a, modes_a, gc, modes_out, modes_b, conj_a=True)
            gb = gb.reshape(b.shape)
        return ga, gb, None, None, None, None


def contract(a: torch.Tensor, modes_a: Sequence[Mode], b: torch.Tensor, modes_b: Sequence[Mode],
             modes_out: Sequence[Mode], fixed: Optional[Dict[Mode, int]] = None) -> torch.Tensor:  # fmt: skip
    if (a.requires_grad or b.requires_grad) and torch.is_grad_enabled():
        # batch modes make dA a reduction over n only; the rule dA = dC.B^H still holds mode-wise
        return _Contract.apply(a, b, tuple(modes_a), tuple(modes_b), tuple(modes_out), fixed)
    return contract_raw(a, modes_a, b, modes_b, modes_out, fixed=fixed)


def tensordot(a: torch.Tensor, b: torch.Tensor, axes_a: Sequence[int], axes_b: Sequence[int]) -> torch.Tensor:
    """Same contract as numpy/torch tensordot: result axes = free(a) ++ free(b)."""
    ma: List[Mode] = [("a", i) for i in range(a.dim())]
    mb: List[Mode] = [("b", i) for i in range(b.dim())]
    for x, y in zip(axes_a, axes_b):
        mb[y] = ma[x]
    shared = {ma[x] for x in axes_a}
    out = [m for m in ma if m not in shared] + [m for m in mb if m not in shared]
    return contract(a, ma, b, mb, out)


def trace(t: torch.Tensor, i: int, j: int) -> torch.Tensor:
    """Partial trace over axes i, j as a contraction with a 2x2 identity (same kernel)."""
    eye = torch.eye(2, dtype=torch.complex64, device=t.device)
    mt: List[Mode] = [("t", k) for k in range(t.dim())]
    out = [m for k, m in enumerate(mt) if k not in (i, j)]
    return contract(t, mt, eye, [mt[i], mt[j]], out)


# ---------------------------------------------------------------------------------------
def slice_leaf(t: torch.Tensor, modes: Sequence[str], fixed: Dict[str, int]) -> Tuple[torch.Tensor, List[str]]:
    """K7 (`tree.slice_arrays`, tensorcircuit/experimental.py:1007) without copies: pinning a mode
    of an extent-2 tensor is a strided *view* (storage offset + remaining power-of-two strides),
    which the kernel reads in place; autograd sees an ordinary `select`."""
    out_modes: List[str] = []
    for m in modes:
        if m in fixed:
            t = t.select(len(out_modes), int(fixed[m]))
        else:
            out_modes.append(m)
    return t, out_modes


# ---- tree schedules -----------------------------------------------------------------------------
_FUSE_SMALL = 64  # elements: what the streaming kernel keeps in shared memory (csrc/tn_kernels.cu)
_FUSE_BIG = 1 << 16
_schedule_cache: Dict[Any, Any] = {}
fuse_skinny_chains = True
plan_layouts = True  # order every intermediate's modes for its consumer (contracted modes lowest)


def _keep_modes(ta: Sequence[str], tb: Sequence[str], occ: Dict[str, int], out_set: set) -> List[str]:
    """Modes of a pairwise result: kept(left) ++ kept(right); a mode survives when the output or a third
    tensor still carries it (hyper-indices included)."""
    sa, sb = set(ta), set(tb)
    keep = []
    for m in dict.fromkeys(list(ta) + list(tb)):
        if m in out_set or occ[m] - (m in sa) - (m in sb) > 0:
            keep.append(m)
    return keep


def _stream_ok(big: Sequence[str], small: Sequence[str], keep: Sequence[str]) -> bool:
    """Would (big, small) -> keep run on the streaming kernel?  (mirror of launch_contract's rule)"""
    sb, ss, sk = set(big), set(small), set(keep)
    nb = sum(1 for m in keep if m in sb and m in ss)
    nn = sum(1 for m in keep if m in ss and m not in sb)
    nk = sum(1 for m in big if m in ss and m not in sk)
    nm = sum(1 for m in keep if m in sb and m not in ss)
    return nk <= 3 and nn <= 3 and nb + nk + nn <= 6 and nm + nb >= 10 and nn <= nm


def build_schedule(inputs: Sequence[Sequence[str]], output: Sequence[str], path: Sequence[Tuple[int, ...]],
                   sliced: Sequence[str] = (), fuse: Optional[bool] = None) -> List[Tuple[int, int, List[str], List[str], List[str], int]]:  # fmt: skip
    """Linear path -> SSA steps (a, b, modes_a, modes_b, keep, out), cached.

    With `fuse`, chains of skinny absorptions are re-associated: when a large tensor absorbs a small one
    and the result immediately absorbs another small one, the two small tensors are contracted with each
    other first (a few dozen elements) and the large tensor is streamed through HBM ONCE instead of twice
    — the tensor-network form of fusing consecutive gates into one statevector pass.  Any association
    order of a tensor network is valid; kept modes are recomputed from occurrence counts."""
    fuse = fuse_skinny_chains if fuse is None else fuse
    key = (tuple(tuple(t) for t in inputs), tuple(output), tuple(tuple(p) for p in path), tuple(sliced), fuse,
           plan_layouts)
    hit = _schedule_cache.get(key)
    if hit is not None:
        return hit
    sl = set(sliced)
    terms: Dict[int, List[str]] = {i: [m for m in t if m not in sl] for i, t in enumerate(inputs)}
    n = len(inputs)
    # linear -> SSA
    live = list(range(n))
    ssa: List[Tuple[int, int]] = []
    nxt = n
    for step in path:
        if len(step) != 2:
            continue
        i, j = step
        ssa.append((live[i], live[j]))
        for k in sorted((i, j), reverse=True):
            live.pop(k)
        live.append(nxt)
        nxt += 1
    if len(live) != 1:
        raise _lib.EngineError("contraction path does not reduce the network to one tensor")
    out_set = set(output)

    def forward(pairs: Sequence[Tuple[int, int, int]]):
        """Recompute every step's kept modes for an SSA step list [(a, b, out)].

        Layout planning (round 2): the physical layout of an intermediate is free, so its mode list is ordered
        for the step that CONSUMES it — the modes that step contracts come last (= the lowest address bits), in
        one canonical order shared by both operands of that step.  The tensor-core kernel then reads 16
        consecutive k of a row as one contiguous 128-byte run from either operand (tn_gemm_tc.cu: with scattered
        k bits every lane of a gather touched its own cache line and the producers, not the tensor pipe, set
        the pace)."""
        occ: Dict[str, int] = {}
        for t in list(terms.values())[:n]:
            for m in t:
                occ[m] = occ.get(m, 0) + 1
        tm = {i: terms[i] for i in range(n)}
        keeps: List[List[str]] = []
        for si, (a, b, o) in enumerate(pairs):  # pass 1: which modes survive each step (order irrelevant)
            ta, tb = tm[a], tm[b]
            keep = list(output) if si == len(pairs) - 1 else _keep_modes(ta, tb, occ, out_set)
            for m in ta:
                occ[m] -= 1
            for m in tb:
                occ[m] -= 1
            for m in keep:
                occ[m] = occ.get(m, 0) + 1
            tm[o] = keep
            keeps.append(keep)
        consumer = {}
        produced_by = {}
        for si, (a, b, o) in enumerate(pairs):
            consumer[a] = si
            consumer[b] = si
            produced_by[o] = si
        rank = {}
        for t in terms.values():
            for m in t:
                rank.setdefault(m, len(rank))
        tm = {i: terms[i] for i in range(n)}
        steps = []
        for si, (a, b, o) in enumerate(pairs):  # pass 2: layouts
            ta, tb = tm[a], tm[b]
            keep = keeps[si]
            c = consumer.get(o)
            if plan_layouts and c is not None and si != len(pairs) - 1:
                survive = set(keeps[c])
                other = pairs[c][1] if pairs[c][0] == o else pairs[c][0]
                # the partner's mode SET is known from pass 1 even when it is produced later
                shared = set(terms[other]) if other < n else set(keeps[produced_by[other]])
                kmodes = sorted((m for m in keep if m not in survive and m in shared), key=lambda m: rank[m])
                ks = set(kmodes)
                keep = [m for m in keep if m not in ks] + kmodes
            tm[o] = keep
            steps.append((a, b, list(ta), list(tb), keep, o))
        return steps, tm

    pairs = [(a, b, n + k) for k, (a, b) in enumerate(ssa)]
    steps, tm = forward(pairs)
    if fuse and len(pairs) > 2:
        changed = True
        rounds = 0
        while changed and rounds < 8:
            changed = False
            rounds += 1
            producer = {o: k for k, (_, _, o) in enumerate(pairs)}
            consumer: Dict[int, int] = {}
            for k, (a, b, _) in enumerate(pairs):
                consumer[a] = k
                consumer[b] = k
            size = {i: 2 ** len(t) for i, t in tm.items()}
            new_pairs: List[Optional[Tuple[int, int, int]]] = list(pairs)
            extra: Dict[int, List[Tuple[int, int, int]]] = {}
            used = set()
            for k, (a, b, o) in enumerate(pairs):
                if k in used or k == len(pairs) - 1:
                    continue
                big, small = (a, b) if size[a] >= size[b] else (b, a)
                if size[big] < _FUSE_BIG or size[small] > _FUSE_SMALL:
                    continue
                k2 = consumer.get(o)
                if k2 is None or k2 in used or k2 == len(pairs) - 1:
                    continue
                a2, b2, o2 = pairs[k2]
                w = b2 if a2 == o else a2
                if size[w] > _FUSE_SMALL or size[o] < _FUSE_BIG:
                    continue
                if w >= n and producer[w] > k:
                    continue  # the second small tensor does not exist yet when the chain starts
                # would the merged small tensor stay small, and the fused step stay on the streaming kernel?
                occ: Dict[str, int] = {}
                for i2, t in tm.items():
                    pass
                merged = [m for m in dict.fromkeys(tm[small] + tm[w])]
                if 2 ** len(merged) > _FUSE_SMALL:
                    continue
                if not _stream_ok(tm[big], merged, tm[o2]):
                    continue
                s12 = nxt
                nxt += 1
                new_pairs[k] = None
                extra.setdefault(k, []).append((small, w, s12))
                new_pairs[k2] = (big, s12, o2)
                used.update((k, k2))
                changed = True
            if changed:
                rebuilt: List[Tuple[int, int, int]] = []
                for k, pr in enumerate(new_pairs):
                    rebuilt.extend(extra.get(k, []))
                    if pr is not None:
                        rebuilt.append(pr)
                pairs = rebuilt
                steps, tm = forward(pairs)
    if len(_schedule_cache) > 64:
        _schedule_cache.clear()
    _schedule_cache[key] = steps
    return steps


def contract_tree(arrays: Sequence[torch.Tensor], inputs: Sequence[Sequence[str]], output: Sequence[str],
                  path: Sequence[Tuple[int, ...]], fixed: Optional[Dict[str, int]] = None) -> torch.Tensor:  # fmt: skip
    """Execute an opt_einsum-style linear path (pop both operands, append the result —
    consumed the same way at tensorcircuit/cons.py:937-950) on the GPU.

    `fixed` pins sliced indices (leaves carrying them are read through offset views).
    Intermediate layouts are  kept(left) ++ kept(right)  (the convention of
    examples/omeco_ready_wave_benchmark.py:198-263); the last step writes `output` order
    directly, so there is no final transpose pass (K2).  The step list comes from
    `build_schedule` (cached; skinny absorption chains fused).
    """
    fixed = dict(fixed or {})
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in arrays)
    if use_native_plans and len(arrays) > 1 and not needs_grad and len(fixed) <= 62:
        return tree_plan(inputs, output, path, sorted(fixed)).execute(arrays, fixed)
    tens: Dict[int, torch.Tensor] = {}
    for i, (t, modes) in enumerate(zip(arrays, inputs)):
        if fixed:
            t, _ = slice_leaf(t, list(modes), fixed)
        tens[i] = t
    if len(arrays) == 1:
        t = tens[0]
        final = [m for m in inputs[0] if m not in fixed]
        if list(final) != list(output):
            t = t.permute([final.index(m) for m in output])
        return t
    steps = build_schedule(inputs, output, path, sorted(fixed))
    r = None
    for a, b, ta, tb, keep, o in steps:
        r = contract(tens.pop(a), ta, tens.pop(b), tb, keep)
        tens[o] = r
    assert r is not None
    return r


# ------------------------------------------------------------------------------------------------
class TreePlan:
    """One contraction tree as a native plan object (`tcb_tn_plan_*`, include/tcb200.h): the SSA schedule,
    every step's descriptor, a liveness-packed workspace layout and the slice-offset tables of the leaves
    are fixed once; `execute` is ONE C call per slice (no per-step Python, no per-step allocation).
    Forward only — differentiated contractions go through `contract_tree`'s autograd steps."""

    def __init__(self, inputs: Sequence[Sequence[str]], output: Sequence[str], path: Sequence[Tuple[int, ...]],
                 sliced: Sequence[str] = ()) -> None:  # fmt: skip
        import ctypes

        self.sliced = sorted(sliced)
        if len(self.sliced) > 62:
            raise _lib.EngineError("TreePlan: more than 62 sliced indices")
        sl = set(self.sliced)
        steps = build_schedule(inputs, output, path, self.sliced)
        n = len(inputs)
        self.nleaves = n
        self.output = list(output)
        # layouts: leaves are contiguous [2]*rank arrays of ALL their modes; intermediates contiguous in `keep` order
        layout: Dict[int, Dict[str, int]] = {}
        counts, pairs = [], []
        for i, t in enumerate(inputs):
            r = len(t)
            if len(set(t)) != r:
                raise _lib.EngineError("TreePlan: repeated index inside one tensor")
            layout[i] = {m: r - 1 - ax for ax, m in enumerate(t)}
            mine = [(self.sliced.index(m), layout[i][m]) for m in t if m in sl]
            counts.append(len(mine))
            pairs.extend(mine)
        remap = {i: i for i in range(n)}
        ids, descs, out_elems = [], [], []
        for k, (a, b, ta, tb, keep, o) in enumerate(steps):
            if not row_side_first(ta, tb, keep):
                a, b, ta, tb = b, a, tb, ta
            descs.append(make_desc(layout[a], layout[b], ta, tb, keep, sl, False, False))
            remap[o] = n + k
            ids.extend((remap[a], remap[b], n + k))
            layout[o] = {m: len(keep) - 1 - ax for ax, m in enumerate(keep)}
            out_elems.append(1 << len(keep))
        self.nsteps = len(steps)
        if self.nsteps == 0:
            raise _lib.EngineError("TreePlan needs at least one pairwise step")
        arr_desc = (_lib.ContractDesc * self.nsteps)(*descs)
        leaf_elems = (ctypes.c_int64 * n)(*[1 << len(t) for t in inputs])
        step_ids = (ctypes.c_int32 * len(ids))(*ids)
        oe = (ctypes.c_int64 * self.nsteps)(*out_elems)
        cnt = (ctypes.c_int32 * n)(*counts)
        flat = [x for pr in pairs for x in pr]
        prs = (ctypes.c_int32 * max(1, len(flat)))(*flat)
        handle = ctypes.c_void_p()
        _lib.check(_lib.load().tcb_tn_plan_create(n, leaf_elems, self.nsteps, step_ids, arr_desc, oe, len(self.sliced),
                                                  cnt, prs, ctypes.byref(handle)))  # fmt: skip
        self._handle = handle
        self.ws_bytes = int(_lib.load().tcb_tn_plan_workspace_size(handle))
        self._ws: Dict[str, torch.Tensor] = {}

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().tcb_tn_plan_destroy(h)
            except Exception:  # pylint: disable=broad-except  (interpreter shutdown)
                pass
            self._handle = None

    def execute(self, arrays: Sequence[torch.Tensor], fixed: Optional[Dict[str, int]] = None) -> torch.Tensor:
        import ctypes

        fixed = fixed or {}
        if sorted(fixed) != self.sliced:
            raise _lib.EngineError("TreePlan.execute: the pinned indices differ from the plan's sliced indices")
        if len(arrays) != self.nleaves:
            raise _lib.EngineError("TreePlan.execute: wrong number of input tensors")
        keep = []  # (keeps converted leaves alive until the launches are queued)
        ptrs = (ctypes.c_void_p * self.nleaves)()
        dev = None
        for i, t in enumerate(arrays):
            if t.dtype != torch.complex64 or t.is_conj() or not t.is_contiguous():
                t = t.to(torch.complex64).resolve_conj().contiguous()
            _lib.require_cuda(t, "input tensor")
            keep.append(t)
            ptrs[i] = t.data_ptr()
            dev = t.device
        ws = self._ws.get(str(dev))
        if ws is None:
            ws = torch.empty(max(1, self.ws_bytes // 8), dtype=torch.complex64, device=dev)
            self._ws[str(dev)] = ws
        out = torch.empty([2] * len(self.output), dtype=torch.complex64, device=dev)
        bits = 0
        for i, m in enumerate(self.sliced):
            bits |= (int(fixed[m]) & 1) << i
        lib = _lib.load()
        _lib.check(lib.tcb_tn_plan_execute(self._handle, ptrs, bits, out.data_ptr(), ws.data_ptr(), self.ws_bytes,
                                           _lib.stream_ptr()))  # fmt: skip
        _lib.launch_count += self.nsteps
        return out


_tree_plans: Dict[Any, TreePlan] = {}
use_native_plans = True


def tree_plan(inputs: Sequence[Sequence[str]], output: Sequence[str], path: Sequence[Tuple[int, ...]],
              sliced: Sequence[str] = ()) -> TreePlan:  # fmt: skip
    key = (tuple(tuple(t) for t in inputs), tuple(output), tuple(tuple(p) for p in path), tuple(sorted(sliced)),
           fuse_skinny_chains)  # fmt: skip
    tp = _tree_plans.get(key)
    if tp is None:
        tp = TreePlan(inputs, output, path, sliced)
        if len(_tree_plans) > 16:
            _tree_plans.clear()
        _tree_plans[key] = tp
    return tp
