"""
Expectation values on a resident statevector.

The reference evaluates <psi|P|psi> as the network [psi, psi*, P...]
(tensorcircuit/basecircuit.py:393-447, tensorcircuit/circuit.py:899-902): a materialised
conjugate copy, a tensordot per operator and a full inner product.  Here the bra is never
built: Pauli strings are a single reduction over the state (`tcb_sv_expect_z` /
`tcb_sv_expect_pauli`), general operators are applied to one scratch copy and closed with
`tcb_sv_inner`.  All entry points are differentiable (torch convention for complex
cotangents, the reference's tests/test_backends.py:992-996).
"""

from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, gates


def _mask(n: int, axes: Sequence[int]) -> int:
    m = 0
    for a in axes:
        m |= 1 << (n - 1 - a)
    return m


def _apply_dense(state: torch.Tensor, n: int, axes: Sequence[int], mat: torch.Tensor) -> None:
    """In place on one state [2^n] or a contiguous batch [B, 2^n] (the same matrix for every member)."""
    bp = _lib.int_array([n - 1 - a for a in axes])
    m = mat.to(torch.complex64).resolve_conj().contiguous()
    nb = 1 if state.dim() == 1 else int(state.shape[0])
    _lib.call("tcb_sv_apply_dense", state.data_ptr(), n, nb, bp, len(axes), m.data_ptr(), 0, _lib.stream_ptr())


_single_masks: dict = {}


def _single_mask(zm: int, device: torch.device) -> torch.Tensor:
    """A one-element device mask table, uploaded once: a host->device copy of a fresh tensor would make every
    expectation call wait for the circuit's passes (the copy is stream-ordered behind them)."""
    key = (zm, str(device))
    t = _single_masks.get(key)
    if t is None:
        if len(_single_masks) > 4096:
            _single_masks.clear()
        t = _single_masks[key] = torch.tensor([zm], dtype=torch.int64, device=device)
    return t


def _pauli_raw(psi: torch.Tensor, n: int, xs: Sequence[int], ys: Sequence[int], zs: Sequence[int]) -> torch.Tensor:
    """psi [2^n] -> complex64 scalar; a contiguous batch [B, 2^n] -> [B]."""
    _lib.require_cuda(psi, "state")
    psi = psi.resolve_conj().contiguous()
    nb = 1 if psi.dim() == 1 else int(psi.shape[0])
    xm = _mask(n, list(xs) + list(ys))
    zm = _mask(n, list(zs) + list(ys))
    if xm == 0:
        out = torch.zeros(nb, dtype=torch.float64, device=psi.device)  # a Z string: real, one slot per member
        zmask = _single_mask(zm, psi.device)
        _lib.call("tcb_sv_expect_z", psi.data_ptr(), n, nb, zmask.data_ptr(), 1, 0, out.data_ptr(), _lib.stream_ptr())
        res = out.to(torch.complex64)
    else:
        out = torch.zeros(nb, 2, dtype=torch.float64, device=psi.device)
        _lib.call("tcb_sv_expect_pauli", psi.data_ptr(), n, nb, xm, zm, len(ys), 0, out.data_ptr(),
                  _lib.stream_ptr())  # fmt: skip
        res = torch.view_as_complex(out).to(torch.complex64)
    return res if psi.dim() == 2 else res.reshape(())


def _apply_pauli(psi: torch.Tensor, n: int, xs: Sequence[int], ys: Sequence[int], zs: Sequence[int]) -> torch.Tensor:
    phi = psi.resolve_conj().clone()
    for name, sites in (("x", xs), ("y", ys), ("z", zs)):
        for a in sites:
            _apply_dense(phi, n, [a], getattr(gates, name)().tensor.to(psi.device))
    return phi


class _PauliExpect(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, psi: torch.Tensor, n: int, xs: Any, ys: Any, zs: Any) -> torch.Tensor:
        ctx.save_for_backward(psi)
        ctx.meta = (n, tuple(xs), tuple(ys), tuple(zs))
        return _pauli_raw(psi, n, xs, ys, zs)

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):  # type: ignore[override]
        (psi,) = ctx.saved_tensors
        n, xs, ys, zs = ctx.meta
        # s = psi^H P psi, P Hermitian:  grad_psi = conj(g) P psi + g P^H psi = 2 Re(g) P psi
        ppsi = _apply_pauli(psi, n, xs, ys, zs)
        if psi.dim() == 2:
            g = g.reshape(-1, 1)
        return (2.0 * g.real.to(torch.float32)) * ppsi, None, None, None, None


def pauli_expectation(psi: torch.Tensor, n: int, xs: Sequence[int], ys: Sequence[int], zs: Sequence[int]) -> torch.Tensor:
    """<psi| X_xs Y_ys Z_zs |psi> as a complex64 scalar tensor (abstractcircuit.py:1523-1603)."""
    from . import autograd

    if autograd.is_batched(psi):  # under torch.vmap: one launch for the whole batch
        phys, lvl = autograd.unwrap_batched(psi)
        with autograd.outside_vmap():
            p2 = phys.to(torch.complex64).resolve_conj().reshape(phys.shape[0], -1).contiguous()
            if p2.requires_grad and torch.is_grad_enabled():
                out = _PauliExpect.apply(p2, n, tuple(xs), tuple(ys), tuple(zs))
            else:
                out = _pauli_raw(p2, n, xs, ys, zs)
        return autograd.rewrap_batched(out, lvl)
    if psi.requires_grad and torch.is_grad_enabled():
        return _PauliExpect.apply(psi, n, tuple(xs), tuple(ys), tuple(zs))
    return _pauli_raw(psi, n, xs, ys, zs)


_mask_cache: dict = {}


def z_expectations(psi: torch.Tensor, n: int, terms: Sequence[Sequence[int]]) -> torch.Tensor:
    """All Z-string expectations in ONE read of the state: terms[t] = qubits carrying Z.
    Returns float64 [len(terms)].  (Pauli-sum expectation, diagonal part.)"""
    _lib.require_cuda(psi, "state")
    psi = psi.resolve_conj().contiguous()
    key = (n, str(psi.device), tuple(tuple(int(q) for q in t) for t in terms))
    zm = _mask_cache.get(key)
    if zm is None:  # the masks of a Hamiltonian are uploaded once, not once per step
        masks = np.array([_mask(n, t) for t in terms], dtype=np.int64)
        zm = torch.from_numpy(masks).to(psi.device)
        if len(_mask_cache) > 64:
            _mask_cache.clear()
        _mask_cache[key] = zm
    out = torch.zeros(len(terms), dtype=torch.float64, device=psi.device)
    _lib.call("tcb_sv_expect_z", psi.data_ptr(), n, 1, zm.data_ptr(), len(terms), 0, out.data_ptr(),
              _lib.stream_ptr())  # fmt: skip
    return out


class _OperatorExpect(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, psi: torch.Tensor, n: int, axes_list: Any, *mats: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(psi, "state")
        psi_c = psi.resolve_conj().contiguous()
        phi = psi_c.clone()
        for axes, m in zip(axes_list, mats):
            d = 1 << len(axes)
            _apply_dense(phi, n, axes, m.reshape(d, d))
        out = torch.zeros(2, dtype=torch.float64, device=psi.device)
        _lib.call("tcb_sv_inner", psi_c.data_ptr(), phi.data_ptr(), n, 1, out.data_ptr(), _lib.stream_ptr())
        ctx.save_for_backward(psi_c, *mats)
        ctx.meta = (n, axes_list)
        return torch.view_as_complex(out.reshape(1, 2)).reshape(()).to(torch.complex64)

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):  # type: ignore[override]
        psi, *mats = ctx.saved_tensors
        n, axes_list = ctx.meta
        if any(ctx.needs_input_grad[3:]):
            raise _lib.EngineError(
                "gradient with respect to an operator tensor was requested from the expectation reduction kernel, "
                "which treats operators as constants; contract the sandwich with set_contractor('tn') instead "
                "(cons.b200_contractor does this on its own when an operator requires grad)"
            )
        # s = psi^H O psi:  grad_psi = conj(g) O psi + g O^H psi   (operators are constants here)
        o_psi = psi.clone()
        oh_psi = psi.clone()
        for axes, m in zip(axes_list, mats):
            d = 1 << len(axes)
            mm = m.reshape(d, d)
            _apply_dense(o_psi, n, axes, mm)
            _apply_dense(oh_psi, n, axes, mm.conj().transpose(0, 1))
        grad = torch.conj(g).to(torch.complex64) * o_psi + g.to(torch.complex64) * oh_psi
        return (grad, None, None) + tuple(None for _ in mats)


def operator_expectation(psi: torch.Tensor, n: int, ops: Sequence[Tuple[torch.Tensor, Tuple[int, ...]]]) -> torch.Tensor:
    """<psi| prod_j O_j |psi> for arbitrary k-qubit operator tensors [out..., in...] on disjoint axes."""
    axes_list = tuple(tuple(ax) for _, ax in ops)
    mats = [t for t, _ in ops]
    return _OperatorExpect.apply(psi, n, axes_list, *mats)
