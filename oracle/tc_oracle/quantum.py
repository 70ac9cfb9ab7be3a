"""
ORACLE (test infrastructure only) — numpy restatement of the reference's Pauli-string-sum operators
(SURVEY §8f rank 1), the step right after the statevector path in every VQE.

Pinned by the reference's own tests: tests/test_quantum.py:1534-1576 (MVP == dense, empty sum, identity
term), tests/test_templates.py:190-211 (operator_expectation: 0.84147 / gradient 0.54032),
tests/test_templates.py:43-60 (sparse_expectation: 1.0 / gradient 0.0), tests/test_quantum.py:1437-1447
(sum of Z on Hamming-weight-i states = 8 - 2 i) — checked in tests/test_oracle_golden.py.
"""

from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence

import numpy as np

_PAULI = [
    np.eye(2, dtype=np.complex128),
    np.array([[0, 1], [1, 0]], dtype=np.complex128),
    np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    np.array([[1, 0], [0, -1]], dtype=np.complex128),
]


def PauliStringSum2MVP(structures: Sequence[Sequence[int]], weights: Sequence[complex]) -> Callable[[Any], Any]:
    """tensorcircuit/quantum.py:2222-2358: mvp(psi) = sum_t w_t P_t psi without a matrix.  Per term: multiply by
    the (+1, -1) mask on every Z and Y axis (`:2296-2307`), reverse every X and Y axis (`:2284-2290,:2341`),
    scale by w * 1j ** ny (`:2283,:2344`).  0: I, 1: X, 2: Y, 3: Z; axis k = qubit k (big-endian flat index)."""
    if not structures:
        return lambda psi: np.zeros_like(np.asarray(psi))
    n = len(structures[0])

    def mvp(psi: Any) -> Any:
        psi = np.asarray(psi)
        flat = psi.ndim == 1
        t = psi.reshape((2,) * n)
        total = np.zeros_like(t)
        for s, w in zip(structures, weights):
            s_arr = np.asarray(s)
            ny = int(np.sum(s_arr == 2))
            term = t
            for a in np.where((s_arr == 3) | (s_arr == 2))[0]:
                term = term * np.array([1.0, -1.0]).reshape([1] * a + [2] + [1] * (n - a - 1))
            flips = np.where((s_arr == 1) | (s_arr == 2))[0]
            if len(flips):
                sl: List[slice] = [slice(None)] * n
                for k in flips:
                    sl[k] = slice(None, None, -1)
                term = term[tuple(sl)]
            total = total + term * (w * (1j) ** ny)
        return total.reshape(-1) if flat else total

    return mvp


def PauliStringSum2Dense(ls: Sequence[Sequence[int]], weight: Optional[Sequence[complex]] = None) -> np.ndarray:
    """tensorcircuit/quantum.py:2361-2387 (the matrix the COO builder `:2390-2456` holds): sum_t w_t kron_k P."""
    n = len(ls[0])
    if weight is None:
        weight = [1.0] * len(ls)
    h = np.zeros((1 << n, 1 << n), dtype=np.complex128)
    for s, w in zip(ls, weight):
        m = np.ones((1, 1), dtype=np.complex128)
        for k in s:
            m = np.kron(m, _PAULI[int(k)])
        h += w * m
    return h


def operator_expectation(c: Any, hamiltonian: np.ndarray) -> float:
    """tensorcircuit/templates/measurements.py:156-191 (dense and sparse branches): Re <psi| H |psi>."""
    w = np.asarray(c.wavefunction()).reshape(-1)
    return float(np.real(np.vdot(w, np.asarray(hamiltonian) @ w)))


sparse_expectation = operator_expectation


def heisenberg_hamiltonian_terms(edges: Sequence[Sequence[int]], n: int, hzz: float = 1.0, hxx: float = 1.0,
                                 hyy: float = 1.0, hz: float = 0.0, hx: float = 0.0, hy: float = 0.0):  # fmt: skip
    """tensorcircuit/quantum.py `heisenberg_hamiltonian`: (structures, weights) of
    sum_edges (hzz ZZ + hxx XX + hyy YY) + sum_nodes (hz Z + hx X + hy Y), zero weights dropped."""
    ls: List[List[int]] = []
    ws: List[float] = []
    for a, b in edges:
        for code, w in ((3, hzz), (1, hxx), (2, hyy)):
            if w != 0:
                s = [0] * n
                s[a] = s[b] = code
                ls.append(s)
                ws.append(w)
    for q in range(n):
        for code, w in ((3, hz), (1, hx), (2, hy)):
            if w != 0:
                s = [0] * n
                s[q] = code
                ls.append(s)
                ws.append(w)
    return ls, ws


# -- sampling ----------------------------------------------------------------------
def probability_sample(shots: int, p: Any, status: Any) -> np.ndarray:
    """tensorcircuit/backends/abstract_backend.py:1828-1861: p /= sum(p); cumsum; r = cuml[-1] (1 - status);
    searchsorted (left).  float64 here: the boundaries of the reference's float32 cumsum are not reproducible."""
    p = np.asarray(p, dtype=np.float64)
    p = p / np.sum(p)
    cuml = np.cumsum(p)
    r = cuml[-1] * (1.0 - np.asarray(status, dtype=np.float64))
    return np.searchsorted(cuml, r)


def measure(psi: Any, index: Sequence[int], status: Sequence[float]):
    """tensorcircuit/basecircuit.py:461-558 (d = 2 branch) on an explicit state: qubit index[k] reads 1 iff
    status[k] - P(0 | earlier outcomes) + 0.31415926e-12 > 0 (`:516-531`); returns (bits, probability)."""
    psi = np.asarray(psi).reshape(-1)
    n = int(round(np.log2(psi.size)))
    t = (np.abs(psi.astype(np.complex128)) ** 2).reshape([2] * n)
    t = t / t.sum()
    bits: List[float] = []
    prob = 1.0
    fixed = {}
    for k, j in enumerate(index):
        sl = tuple(fixed.get(q, slice(None)) for q in range(n))
        cond = t[sl]
        axes = [q for q in range(n) if q not in fixed]
        ax = axes.index(j)
        m = cond.sum(axis=tuple(a for a in range(cond.ndim) if a != ax))
        pu = m[0] / m.sum() if m.sum() > 0 else 1.0
        one = status[k] - pu + 0.31415926e-12 > 0
        bits.append(1.0 if one else 0.0)
        prob *= (1.0 - pu) if one else pu
        fixed[j] = 1 if one else 0
    return np.array(bits), prob


def sample_int2bin(sample: Any, n: int) -> np.ndarray:  # tensorcircuit/quantum.py:3587-3605
    sample = np.asarray(sample)
    return (sample[..., None] >> np.arange(n)[::-1]) % 2
