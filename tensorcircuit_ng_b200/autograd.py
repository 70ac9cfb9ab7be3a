"""
Differentiable statevector evolution (north-star kernel (3)).

`evolve(cc, gatebuf, init)` is a `torch.autograd.Function` whose forward runs the fused
pass programs and whose backward is the *adjoint method*: instead of saving every
intermediate 2^n state (what autograd over `torch.tensordot` does in the reference,
tensorcircuit/backends/pytorch_backend.py:775-786), it keeps only the final state and
walks the gates in reverse, un-computing |psi> with U^dagger while propagating the
cotangent |lam>, and reducing  dL/dU = sum lam (x) conj(psi_in)  per gate with
`tcb_sv_gate_grad`.  Memory: 2 states + 1 scratch, independent of depth.

Valid for unitary gates (every factory in gates.py except user matrices passed to
`any` / `diagonal`); `assume_unitary = False` switches to recomputing psi_in from the
start for each gate (O(G^2) passes, exact for arbitrary matrices).
"""

from __future__ import annotations

from typing import Any, List, Optional

import torch

from . import _lib, svengine
from .passplan import GateOp

assume_unitary = True


def _forward(cc: "svengine.CompiledCircuit", gatebuf: torch.Tensor, init: Optional[torch.Tensor]) -> torch.Tensor:
    nbits = cc.plan.nbits
    if init is None:
        state = torch.empty(1 << nbits, dtype=torch.complex64, device=gatebuf.device)
        _lib.require_cuda(state, "state")
        cc.start(state, gatebuf)
    else:
        if cc.prefix_levels:
            raise _lib.EngineError("a circuit compiled with absorbed leading gates cannot start from `inputs`")
        state = init.detach().to(torch.complex64).resolve_conj().reshape(-1).clone()
        _lib.require_cuda(state, "inputs")
    cc.run(state, gatebuf)
    return state


def _apply_single(state: torch.Tensor, nbits: int, nq: int, op: GateOp, mat: torch.Tensor) -> None:
    """Apply one k-qubit dense matrix (device tensor, row-major) in place (unfused launch)."""
    bp = _lib.int_array([nq - 1 - q for q in op.qubits])
    _lib.call("tcb_sv_apply_dense", state.data_ptr(), nbits, 1, bp, op.k, mat.data_ptr(), 0, _lib.stream_ptr())


def _dense_matrix(gatebuf: torch.Tensor, op: GateOp) -> torch.Tensor:
    d = 1 << op.k
    if op.kind[0] == "diagvec":
        return torch.diag(gatebuf[op.mat_off : op.mat_off + d])
    return gatebuf[op.mat_off : op.mat_off + d * d].reshape(d, d)


class _Evolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, gatebuf: torch.Tensor, init: Optional[torch.Tensor], cc: Any) -> torch.Tensor:
        state = _forward(cc, gatebuf.detach(), init)
        ctx.cc = cc
        ctx.has_init = init is not None
        ctx.save_for_backward(gatebuf.detach(), state)
        return state

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore[override]
        gatebuf, psi_out = ctx.saved_tensors
        cc = ctx.cc
        nbits = cc.plan.nbits
        nq = nbits
        lam = grad_out.to(torch.complex64).resolve_conj().reshape(-1).clone()
        psi = psi_out.clone()
        grad_buf = torch.zeros_like(gatebuf)
        stream_ops: List[GateOp] = cc.ops
        for op in reversed(stream_ops):
            d = 1 << op.k
            u = _dense_matrix(gatebuf, op)
            udag = u.conj().transpose(0, 1).contiguous()
            if not assume_unitary:
                raise _lib.EngineError(
                    "autograd.assume_unitary=False (recompute mode) is not implemented in this round"
                )
            # psi_in = U^dagger psi_out
            _apply_single(psi, nbits, nq, op, udag)
            if op.k > 2:
                raise _lib.EngineError(
                    f"gradient through a {op.k}-qubit gate is not supported yet (tcb_sv_gate_grad: k <= 2)"
                )
            g = torch.zeros(d * d * 2, dtype=torch.float64, device=gatebuf.device)
            bp = _lib.int_array([nq - 1 - q for q in op.qubits])
            _lib.call("tcb_sv_gate_grad", lam.data_ptr(), psi.data_ptr(), nbits, 1, bp, op.k, g.data_ptr(), 0,
                      _lib.stream_ptr())  # fmt: skip
            gc = torch.view_as_complex(g.reshape(d * d, 2)).to(torch.complex64)
            if op.kind[0] == "diagvec":
                grad_buf[op.mat_off : op.mat_off + d] = gc.reshape(d, d).diagonal()
            else:
                grad_buf[op.mat_off : op.mat_off + d * d] = gc
            # lam_in = U^dagger lam_out
            _apply_single(lam, nbits, nq, op, udag)
        grad_init = lam if ctx.has_init and ctx.needs_input_grad[1] else None
        return grad_buf, grad_init, None


def evolve(cc: "svengine.CompiledCircuit", gatebuf: torch.Tensor, init: Optional[torch.Tensor]) -> torch.Tensor:
    needs = torch.is_grad_enabled() and (gatebuf.requires_grad or (init is not None and init.requires_grad))
    if not needs:
        return _forward(cc, gatebuf, init)
    return _Evolve.apply(gatebuf, init, cc)
