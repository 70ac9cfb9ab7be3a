"""
Hamiltonian expectation on the circuit's output state (`tensorcircuit/templates/measurements.py:156-216`).

`operator_expectation(c, hamiltonian)` takes what the reference takes: a dense matrix, or the "sparse"
Hamiltonian — which in this engine is the matrix-free `quantum.PauliStringSum` (what
`PauliStringSum2COO` / `heisenberg_hamiltonian(sparse=True)` return).  The sparse branch is ONE
`tcb_sv_pauli_sum` launch over the resident state; its backward is one more (H psi), feeding the
adjoint walk of `autograd.py`.
"""

from __future__ import annotations

from typing import Any

import torch

from ..quantum import PauliStringSum


def sparse_expectation(c: Any, hamiltonian: PauliStringSum) -> torch.Tensor:
    """`measurements.py:178-191`: Re <psi| H |psi>, H kept matrix-free."""
    if not isinstance(hamiltonian, PauliStringSum):
        raise TypeError("sparse_expectation expects a quantum.PauliStringSum (PauliStringSum2COO(...))")
    psi = c.wavefunction()
    return hamiltonian.expectation(psi.reshape(-1)).real


def operator_expectation(c: Any, hamiltonian: Any) -> torch.Tensor:
    """`measurements.py:156-175`: dense matrix or matrix-free Pauli sum; a real scalar tensor."""
    if isinstance(hamiltonian, PauliStringSum):
        return sparse_expectation(c, hamiltonian)
    w = c.wavefunction().reshape(-1, 1)
    h = hamiltonian if isinstance(hamiltonian, torch.Tensor) else torch.as_tensor(hamiltonian)
    h = h.to(device=w.device, dtype=w.dtype)
    return (w.conj().transpose(0, 1) @ h @ w)[0, 0].real
