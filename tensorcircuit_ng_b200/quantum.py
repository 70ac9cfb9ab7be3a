"""
Pauli-string-sum operators on a resident statevector, matrix-free.

The reference turns H = sum_t w_t P_t into a COO matrix (`tensorcircuit/quantum.py:2390-2456`) and
evaluates <psi|H|psi> with a sparse matvec (`tensorcircuit/templates/measurements.py:156-191`), or loops
over the terms with flips and broadcast masks (`PauliStringSum2MVP`, `:2222-2358`).  Here a sum is three
device arrays (flip mask, sign mask, coefficient per term, sorted by flip mask) and ONE kernel
(`tcb_sv_pauli_sum`) that produces H psi and / or <psi|H|psi>: no 2^n x 2^n object exists in any form, and
every distinct flip pattern costs one read of the state.

`PauliStringSum2COO` keeps the reference's name and call signature but returns a `PauliStringSum`
(the engine's stand-in for the sparse matrix: `operator_expectation` / `sparse_expectation` accept it
where the reference accepts the COO tensor); `numpy=True` returns the scipy COO matrix itself.
"""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_PAULI_NP = [
    np.eye(2, dtype=np.complex64),
    np.array([[0, 1], [1, 0]], dtype=np.complex64),
    np.array([[0, -1j], [1j, 0]], dtype=np.complex64),
    np.array([[1, 0], [0, -1]], dtype=np.complex64),
]


def _to_host(x: Any) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class PauliStringSum:
    """H = sum_t w_t P_t over n qubits; structures[t][k] in {0: I, 1: X, 2: Y, 3: Z} acts on qubit k.
    Terms with equal Pauli strings are merged, zero weights dropped, the rest sorted by flip mask."""

    def __init__(self, structures: Sequence[Sequence[int]], weights: Optional[Sequence[complex]] = None,
                 nqubits: Optional[int] = None) -> None:  # fmt: skip
        ls = _to_host(structures).astype(np.int64).reshape(len(structures), -1) if len(structures) else None
        # trainable weights (a torch tensor in an autograd graph): the device tables below hold detached values, so
        # `expectation` goes term by term (dE/dw_t = <P_t>) and `mvp` refuses instead of dropping the gradient
        self._live_weights: Optional[torch.Tensor] = None
        if isinstance(weights, torch.Tensor) and (weights.requires_grad or weights.grad_fn is not None):
            self._live_weights = weights
            self._structures = ls
        if ls is None and nqubits is None:
            raise ValueError("an empty Pauli sum needs `nqubits`")
        self.n = int(ls.shape[1]) if ls is not None else int(nqubits)  # type: ignore[union-attr]
        if self.n > 63:
            raise ValueError("PauliStringSum supports up to 63 qubits")
        w = np.ones(len(structures), dtype=np.complex128) if weights is None else _to_host(weights).astype(np.complex128)
        if ls is not None and ((ls < 0) | (ls > 3)).any():
            raise ValueError("Pauli codes must be 0 (I), 1 (X), 2 (Y) or 3 (Z)")
        merged: Dict[Tuple[int, int], complex] = {}
        for t in range(len(structures)):
            x = z = ny = 0
            for k in range(self.n):
                code = int(ls[t, k])  # type: ignore[index]
                bit = 1 << (self.n - 1 - k)  # qubit 0 = most significant bit of the flat index
                if code in (1, 2):
                    x |= bit
                if code in (2, 3):
                    z |= bit
                if code == 2:
                    ny += 1
            merged[(x, z)] = merged.get((x, z), 0.0) + complex(w[t]) * (1j) ** ny
        items = sorted(((x, z, c) for (x, z), c in merged.items() if c != 0), key=lambda it: (it[0], it[1]))
        self.xmask = np.array([it[0] for it in items], dtype=np.uint64)
        self.zmask = np.array([it[1] for it in items], dtype=np.uint64)
        self.coef = np.array([it[2] for it in items], dtype=np.complex64)
        # c_t carries i^ny; the WEIGHT w_t = c_t / i^ny is real for a Hermitian sum
        ny_of = np.array([bin(int(x) & int(z)).count("1") for x, z in zip(self.xmask, self.zmask)], dtype=np.int64)
        wts = self.coef.astype(np.complex128) * (-1j) ** ny_of if len(items) else np.zeros(0, np.complex128)
        self.hermitian = bool(np.all(np.abs(wts.imag) <= 1e-7 * np.maximum(1.0, np.abs(wts.real))))
        self._ny = ny_of
        self._dev: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._adjoint: Optional["PauliStringSum"] = None

    # -- bookkeeping ------------------------------------------------------------------------------
    @property
    def nterms(self) -> int:
        return int(len(self.coef))

    @property
    def shape(self) -> Tuple[int, int]:
        return (1 << self.n, 1 << self.n)

    def adjoint(self) -> "PauliStringSum":
        """H^dagger = sum conj(w_t) P_t (Pauli strings are Hermitian)."""
        if self.hermitian:
            return self
        if self._adjoint is None:
            h = PauliStringSum.__new__(PauliStringSum)
            h.n, h.xmask, h.zmask, h._ny = self.n, self.xmask, self.zmask, self._ny
            wts = self.coef.astype(np.complex128) * (-1j) ** self._ny
            h.coef = (np.conj(wts) * (1j) ** self._ny).astype(np.complex64)
            h.hermitian, h._dev, h._adjoint = False, {}, self
            self._adjoint = h
        return self._adjoint

    def _tables(self, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        key = str(device)
        t = self._dev.get(key)
        if t is None:  # uploaded once per device, reused by every step
            t = (torch.from_numpy(self.xmask.view(np.int64).copy()).to(device),
                 torch.from_numpy(self.zmask.view(np.int64).copy()).to(device),
                 torch.from_numpy(self.coef.copy()).to(device))  # fmt: skip
            self._dev[key] = t
        return t

    def _check(self, psi: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(psi, "state")
        if psi.numel() != (1 << self.n):
            raise ValueError(f"state has {psi.numel()} amplitudes, the Pauli sum acts on {self.n} qubits")
        return psi.to(torch.complex64).resolve_conj().contiguous().reshape(-1)

    # -- raw launches -----------------------------------------------------------------------------
    def _launch(self, psi: torch.Tensor, want_state: bool, want_value: bool) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:  # fmt: skip
        """psi: [2^n] or a batch [B, 2^n] (contiguous); value: float64 [2] or [B, 2]."""
        xs, zs, cs = self._tables(psi.device)
        nb = 1 if psi.dim() == 1 else int(psi.shape[0])
        out = torch.empty_like(psi) if want_state else None
        val = torch.zeros((2,) if psi.dim() == 1 else (nb, 2), dtype=torch.float64, device=psi.device) if want_value else None
        _lib.call("tcb_sv_pauli_sum", psi.data_ptr(), self.n, nb, xs.data_ptr(), zs.data_ptr(), cs.data_ptr(),
                  self.nterms, 0, out.data_ptr() if out is not None else None, 0,
                  val.data_ptr() if val is not None else None, _lib.stream_ptr())  # fmt: skip
        return out, val

    # -- differentiable entry points --------------------------------------------------------------
    def mvp(self, psi: torch.Tensor) -> torch.Tensor:
        """H psi, same shape as `psi` (flat [2^n] or [2]*n)."""
        shape = psi.shape
        if self._live_weights is not None and torch.is_grad_enabled():
            raise _lib.EngineError("PauliStringSum.mvp does not differentiate its weights; use `expectation`, or "
                                   "build the sum from detached weights")
        return _MVP.apply(self._check(psi), self).reshape(shape)

    __call__ = mvp

    def expectation(self, psi: torch.Tensor) -> torch.Tensor:
        """<psi|H|psi> as a complex64 scalar (real up to rounding for a Hermitian sum)."""
        from . import autograd, expect

        if self._live_weights is not None and torch.is_grad_enabled():
            if autograd.is_batched(psi):
                raise _lib.EngineError("trainable Pauli-sum weights are not supported under vmap")
            flat = self._check(psi)
            w = self._live_weights.to(flat.device).reshape(-1)
            total = torch.zeros((), dtype=torch.complex64, device=flat.device)
            for t in range(self._structures.shape[0]):  # type: ignore[union-attr]
                row = self._structures[t]  # type: ignore[index]
                xs, ys, zs = ([int(k) for k in np.nonzero(row == c)[0]] for c in (1, 2, 3))
                total = total + w[t].to(torch.complex64) * expect.pauli_expectation(flat, self.n, xs, ys, zs)
            return total
        if autograd.is_batched(psi):  # under torch.vmap: one launch for the whole batch
            phys, lvl = autograd.unwrap_batched(psi)
            with autograd.outside_vmap():
                p2 = phys.to(torch.complex64).resolve_conj().reshape(phys.shape[0], -1).contiguous()
                if p2.shape[1] != (1 << self.n):
                    raise ValueError(f"state has {p2.shape[1]} amplitudes, the Pauli sum acts on {self.n} qubits")
                out = _Expect.apply(p2, self)
            return autograd.rewrap_batched(out, lvl)
        return _Expect.apply(self._check(psi), self)

    # -- small-n conversions (tests, interop) -----------------------------------------------------
    def to_dense_numpy(self) -> np.ndarray:
        if self.n > 14:
            raise ValueError("dense form of a Pauli sum is only offered up to 14 qubits")
        dim = 1 << self.n
        h = np.zeros((dim, dim), dtype=np.complex128)
        rows = np.arange(dim, dtype=np.uint64)
        for x, z, c in zip(self.xmask, self.zmask, self.coef):
            cols = rows ^ x
            par = np.array([bin(int(v)).count("1") & 1 for v in (cols & z)])
            h[rows.astype(np.int64), cols.astype(np.int64)] += complex(c) * (1 - 2 * par)
        return h

    def to_coo_numpy(self) -> Any:
        import scipy.sparse as sp

        return sp.coo_matrix(self.to_dense_numpy())


class _MVP(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, psi: torch.Tensor, h: PauliStringSum) -> torch.Tensor:
        ctx.h = h
        out, _ = h._launch(psi, True, False)
        return out

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):  # type: ignore[override]
        # y = H psi (holomorphic, linear): torch's cotangent of psi is H^dagger g
        gg = g.to(torch.complex64).resolve_conj().contiguous().reshape(-1)
        out, _ = ctx.h.adjoint()._launch(gg, True, False)
        return out, None


class _Expect(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, psi: torch.Tensor, h: PauliStringSum) -> torch.Tensor:
        ctx.h = h
        ctx.save_for_backward(psi)
        _, val = h._launch(psi, False, True)
        if psi.dim() == 2:
            return torch.view_as_complex(val).to(torch.complex64)
        return torch.view_as_complex(val.reshape(1, 2)).reshape(()).to(torch.complex64)

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):  # type: ignore[override]
        (psi,) = ctx.saved_tensors
        h: PauliStringSum = ctx.h
        # s = psi^H H psi:  grad_psi = conj(g) H psi + g H^dagger psi   (= 2 Re(g) H psi when H is Hermitian)
        hp, _ = h._launch(psi, True, False)
        if psi.dim() == 2:
            g = g.reshape(-1, 1)
        if h.hermitian:
            return (2.0 * g.real.to(torch.float32)) * hp, None
        hdp, _ = h.adjoint()._launch(psi, True, False)
        return torch.conj(g).to(torch.complex64) * hp + g.to(torch.complex64) * hdp, None


# -- the reference's names --------------------------------------------------------------------------
def PauliStringSum2MVP(structures: Sequence[Sequence[int]], weights: Sequence[complex]) -> Callable[[torch.Tensor], torch.Tensor]:
    """`tensorcircuit/quantum.py:2222-2358`: returns mvp(psi) = sum_t w_t P_t psi for psi of shape [2^n]
    or [2]*n; an empty sum maps to zeros_like(psi)."""
    if len(structures) == 0:
        return lambda psi: torch.zeros_like(psi)
    return PauliStringSum(structures, weights).mvp


def PauliStringSum2COO(ls: Sequence[Sequence[int]], weight: Optional[Sequence[complex]] = None, numpy: bool = False) -> Any:
    """`tensorcircuit/quantum.py:2390-2456`.  Returns the matrix-free `PauliStringSum` (accepted wherever the
    reference takes the sparse Hamiltonian); `numpy=True` gives the scipy COO matrix (small n)."""
    h = PauliStringSum(ls, weight)
    return h.to_coo_numpy() if numpy else h


def PauliString2COO(l: Sequence[int], weight: Optional[complex] = None) -> PauliStringSum:
    """`tensorcircuit/quantum.py` PauliString2COO: a single string."""
    return PauliStringSum([list(_to_host(l))], None if weight is None else [weight])


def PauliStringSum2Dense(ls: Sequence[Sequence[int]], weight: Optional[Sequence[complex]] = None, numpy: bool = False) -> Any:
    """`tensorcircuit/quantum.py:2361-2387`: the dense matrix (small n only)."""
    m = PauliStringSum(ls, weight).to_dense_numpy()
    if numpy:
        return m
    t = torch.from_numpy(m.astype(np.complex64))
    return t.cuda() if torch.cuda.is_available() else t


def heisenberg_hamiltonian(g: Any, hzz: float = 1.0, hxx: float = 1.0, hyy: float = 1.0, hz: float = 0.0,
                           hx: float = 0.0, hy: float = 0.0, sparse: bool = True, numpy: bool = False) -> Any:  # fmt: skip
    """`tensorcircuit/quantum.py:2131-2220`: sum over edges of hzz ZZ + hxx XX + hyy YY plus fields on every
    node, in the reference's term order; `sparse=True` returns the matrix-free `PauliStringSum`."""
    n = len(g.nodes)
    ls: List[List[int]] = []
    ws: List[float] = []
    for e in g.edges:
        for code, w in ((3, hzz), (1, hxx), (2, hyy)):
            if w != 0:
                r = [0] * n
                r[e[0]] = r[e[1]] = code
                ls.append(r)
                ws.append(w)
    for node in g.nodes:
        for code, w in ((3, hz), (1, hx), (2, hy)):
            if w != 0:
                r = [0] * n
                r[node] = code
                ls.append(r)
                ws.append(w)
    if sparse:
        return PauliStringSum2COO(ls, ws, numpy=numpy)
    return PauliStringSum2Dense(ls, ws, numpy=numpy)


# -- measurement result formats (`tensorcircuit/quantum.py:3587-3902`) -------------------------------
def sample_int2bin(sample: torch.Tensor, n: int, dim: Optional[int] = None) -> torch.Tensor:
    """[shots] flat indices -> [shots, n] bits, qubit 0 first (`:3587-3613`)."""
    if dim not in (None, 2):
        raise NotImplementedError("qudits are outside the B200 hot-path scope (SURVEY §2.1)")
    shifts = torch.arange(n - 1, -1, -1, device=sample.device, dtype=torch.int64)
    return (sample.to(torch.int64)[..., None] >> shifts) & 1


def sample_bin2int(sample: torch.Tensor, n: int, dim: Optional[int] = None) -> torch.Tensor:
    """[shots, n] bits -> [shots] flat indices."""
    if dim not in (None, 2):
        raise NotImplementedError("qudits are outside the B200 hot-path scope (SURVEY §2.1)")
    w = 2 ** torch.arange(n - 1, -1, -1, device=sample.device, dtype=torch.int64)
    return (sample.to(torch.int64) * w).sum(-1)


def sample2count(sample: torch.Tensor, n: int, jittable: bool = True, dim: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """count_tuple: (distinct flat indices, their counts)."""
    return torch.unique(sample.to(torch.int64), return_counts=True)


def count_s2d(srepr: Tuple[torch.Tensor, torch.Tensor], n: int, dim: Optional[int] = None) -> torch.Tensor:
    """count_tuple -> count_vector of length 2^n."""
    idx, cnt = srepr
    out = torch.zeros(1 << n, dtype=cnt.dtype, device=cnt.device)
    out[idx] = cnt
    return out


def count_tuple2dict(count_tuple: Tuple[torch.Tensor, torch.Tensor], n: int, key: str = "bin", dim: Optional[int] = None) -> Dict[Any, int]:
    idx, cnt = (t.cpu().tolist() for t in count_tuple)
    if key == "int":
        return {int(i): int(c) for i, c in zip(idx, cnt)}
    return {format(int(i), f"0{n}b"): int(c) for i, c in zip(idx, cnt)}


def sample2all(sample: torch.Tensor, n: int, format: str = "count_vector", jittable: bool = False, dim: Optional[int] = None) -> Any:
    """`:3840-3902`: sample_int / sample_bin / count_vector / count_tuple / count_dict_bin / count_dict_int."""
    if sample.dim() == 1:
        s_int, s_bin = sample, None
    elif sample.dim() == 2:
        s_int, s_bin = sample_bin2int(sample, n), sample
    else:
        raise ValueError("unrecognized tensor shape for sample")
    if format == "sample_int":
        return s_int
    if format == "sample_bin":
        return s_bin if s_bin is not None else sample_int2bin(s_int, n)
    ct = sample2count(s_int, n, jittable=jittable)
    if format == "count_tuple":
        return ct
    if format == "count_vector":
        return count_s2d(ct, n)
    if format == "count_dict_bin":
        return count_tuple2dict(ct, n, key="bin")
    if format == "count_dict_int":
        return count_tuple2dict(ct, n, key="int")
    raise ValueError("unsupported format %s for finite shots measurement" % format)
