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
    out_set = set(modes_out)
    free_a = sum(1 for m in modes_a if m in out_set and m not in modes_b)
    free_b = sum(1 for m in modes_b if m in out_set and m not in modes_a)
    if free_b > free_a:
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
    pc = {m: n_out - 1 - i for i, m in enumerate(modes_out)}
    a_off = sum(fixed[m] << p for m, p in pa.items() if m in fixed)
    b_off = sum(fixed[m] << p for m, p in pb.items() if m in fixed)
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
    d.conj_a = int(bool(conj_a) ^ bool(cja))
    d.conj_b = int(bool(conj_b) ^ bool(cjb))
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


def contract_tree(arrays: Sequence[torch.Tensor], inputs: Sequence[Sequence[str]], output: Sequence[str],
                  path: Sequence[Tuple[int, ...]], fixed: Optional[Dict[str, int]] = None) -> torch.Tensor:  # fmt: skip
    """Execute an opt_einsum-style linear path (pop both operands, append the result —
    consumed the same way at tensorcircuit/cons.py:937-950) on the GPU.

    `fixed` pins sliced indices (leaves carrying them are read through offset views).
    Intermediate layouts are  kept(left) ++ kept(right)  (the convention of
    examples/omeco_ready_wave_benchmark.py:198-263); the last step writes `output` order
    directly, so there is no final transpose pass (K2).
    """
    fixed = dict(fixed or {})
    terms: List[List[str]] = []
    tens: List[torch.Tensor] = []
    for t, modes in zip(arrays, inputs):
        if fixed:
            t, modes = slice_leaf(t, list(modes), fixed)
        terms.append(list(modes))
        tens.append(t)
    out_set = set(output)
    steps = [p for p in path if len(p) == 2]
    for si, (i, j) in enumerate(steps):
        a, b = tens[i], tens[j]
        ta, tb = terms[i], terms[j]
        rest = set(out_set)
        for k, t in enumerate(terms):
            if k not in (i, j):
                rest.update(t)
        last = si == len(steps) - 1
        if last:
            keep = [m for m in output]
        else:
            keep = [m for m in dict.fromkeys(ta + tb) if m in rest]
        r = contract(a, ta, b, tb, keep)
        for k in sorted((i, j), reverse=True):
            terms.pop(k)
            tens.pop(k)
        terms.append(list(keep))
        tens.append(r)
    if len(tens) != 1:
        raise _lib.EngineError("contraction path does not reduce the network to one tensor")
    final, t = terms[0], tens[0]
    if list(final) != list(output):
        t = t.permute([final.index(m) for m in output])
    return t
