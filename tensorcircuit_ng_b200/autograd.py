"""
Differentiable statevector evolution (north-star kernel (3)).

`evolve(cc, gatebuf, init)` is a `torch.autograd.Function` whose forward runs the fused pass programs and whose
backward is the *adjoint method*: instead of saving every intermediate 2^n state (what autograd over
`torch.tensordot` does in the reference, tensorcircuit/backends/pytorch_backend.py:775-786), it keeps only the
final state and walks the circuit in reverse, un-computing |psi> with U^dagger while propagating the cotangent
|lam> and reducing  dL/dU = sum lam (x) conj(psi_in).  Memory: 2 states + 1 scratch, independent of depth.

The walk is layered (`_AdjointTables.segments`): runs of diagonal gates, runs of one-qubit gates on distinct
qubits and runs of constant gates are differentiated / un-applied as a whole (`_DiagRun`, `_OneQubitRun`,
`_ConstRun`, `_FusedUnapply`: a few reads of psi and lam + fused sub-circuits), everything else takes one fused
launch per gate (`tcb_sv_adjoint_step`).  Under `torch.vmap` (`backend.vmap / vvag`) the same walk runs over a
batch of states at once (`_EvolveBatched`).

Valid for unitary gates (every factory in gates.py except user matrices passed to `any` / `diagonal`);
`assume_unitary = False` (recompute psi_in from the start for each gate, exact for arbitrary matrices) is not
implemented.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from . import _lib, svengine

assume_unitary = True


def _forward(cc: "svengine.CompiledCircuit", gatebuf: torch.Tensor, init: Optional[torch.Tensor]) -> torch.Tensor:
    nbits = cc.plan.nbits
    if init is None:
        state = torch.empty(1 << nbits, dtype=torch.complex64, device=gatebuf.device)
        _lib.require_cuda(state, "state")
        cc.start_and_run(state, gatebuf)
        return state
    else:
        if cc.prefix_levels:
            raise _lib.EngineError("a circuit compiled with absorbed leading gates cannot start from `inputs`")
        state = init.detach().to(torch.complex64).resolve_conj().reshape(-1).clone()
        _lib.require_cuda(state, "inputs")
    cc.run(state, gatebuf)
    return state


diag_run_min = 3  # shorter runs of diagonal gates are walked gate by gate
one_qubit_run_min = 4  # likewise for runs of one-qubit gates on distinct qubits
const_run_min = 2  # ... and for runs of constant gates (un-applied together, no gradients)
layered_adjoint = True


class _DiagRun:
    """A run [first, last) of consecutive diagonal gates of the circuit (program order) for the backward walk."""

    def __init__(self, cc: "svengine.CompiledCircuit", first: int, last: int, dense_offs: List[int],
                 device: torch.device) -> None:  # fmt: skip
        nq = cc.plan.nbits
        ops = cc.ops[first:last]
        self.first, self.last = first, last
        bits: List[int] = []
        gidx: List[int] = []   # per (gate, c): index of U^dagger[c, c] in the dense dag buffer
        ok: List[bool] = []
        for op, off in zip(ops, dense_offs):
            d = 1 << op.k
            bp = [nq - 1 - q for q in op.qubits]
            bits += [bp[0], bp[1] if op.k == 2 else -1]
            for c in range(4):
                ok.append(c < d)
                gidx.append(off + c * (d + 1) if c < d else 0)
        self.ngates = len(ops)
        self.gate_bits = _lib.int_array(bits)
        self.gidx = torch.tensor(gidx, dtype=torch.long, device=device)
        self.ok = torch.tensor(ok, dtype=torch.bool, device=device)
        # the daggered run as a circuit of its own: fused passes un-apply it on psi and on lam
        structure = [(op.qubits, ("diag",), (1 << op.k) ** 2) for op in reversed(ops)]
        self.sub = svengine.compile_circuit(nq, structure, device, absorb_prefix=False)
        sub_idx: List[int] = []
        for op, off in zip(reversed(ops), reversed(dense_offs)):
            sub_idx += list(range(off, off + (1 << op.k) ** 2))
        self.sub_idx = torch.tensor(sub_idx, dtype=torch.long, device=device)

    def backward(self, nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor, g_all: torch.Tensor) -> None:
        """lam, psi: [2^n] or a contiguous batch [B, 2^n]; dag [.., T], g_all [.., T, 2] likewise."""
        self.grads(nbits, lam, psi, dag, g_all)
        self.unapply(lam, psi, dag)

    def grads(self, nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor, g_all: torch.Tensor) -> None:
        """From the states AFTER the run."""
        nb = 1 if lam.dim() == 1 else int(lam.shape[0])
        dag2, g2 = dag.reshape(nb, -1), g_all.reshape(nb, -1, 2)
        bins = torch.zeros(nb, self.ngates * 4, 2, dtype=torch.float64, device=psi.device)
        _lib.call("tcb_sv_cross_marginals", lam.data_ptr(), psi.data_ptr(), nbits, nb, self.ngates, self.gate_bits,
                  bins.data_ptr(), self.ngates * 4, _lib.stream_ptr())  # fmt: skip
        # dL/dd_j[c] = d_j[c] * bins_j[c];  d_j[c] = conj(U^dagger[c, c])
        dvals = dag2[:, self.gidx].conj().to(torch.complex128)
        grad = torch.view_as_real(dvals * torch.view_as_complex(bins))
        g2.index_add_(1, self.gidx[self.ok], grad[:, self.ok])

    def unapply(self, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor) -> None:
        nb = 1 if lam.dim() == 1 else int(lam.shape[0])
        sub_gb = dag.reshape(nb, -1)[:, self.sub_idx].contiguous()
        self.sub.run(psi, sub_gb, batch=nb, gate_batch_stride=int(sub_gb.shape[1]))
        self.sub.run(lam, sub_gb, batch=nb, gate_batch_stride=int(sub_gb.shape[1]))


class _OneQubitRun:
    """A run [first, last) of consecutive one-qubit gates on DISTINCT qubits.  They commute and every other gate
    of the run is a unitary on another qubit, so it drops out of the partial trace:
    dL/dU_q = U_q C_q  with  C_q[r][c] = sum_rest lam_0[rest, r] conj(psi_0[rest, c])  of the states BEFORE the
    run.  The run is un-applied as one fused sub-circuit, then all C_q come from ceil((m - 3) / 7) reads of the
    two states (`tcb_sv_cross_rdm`) instead of one adjoint step per gate."""

    def __init__(self, cc: "svengine.CompiledCircuit", first: int, last: int, dense_offs: List[int],
                 device: torch.device, indices: Optional[List[int]] = None) -> None:  # fmt: skip
        """Gates cc.ops[first:last], or the explicit `indices` (one round of a block of one-qubit gates in which
        qubits repeat: gates on different qubits commute, so the block is its rounds one after the other)."""
        nq = cc.plan.nbits
        ops = cc.ops[first:last] if indices is None else [cc.ops[i] for i in indices]
        self.first, self.last = (first, last) if indices is None else (-1, -2)
        low = min(3, nq)
        bitpos = [nq - 1 - op.qubits[0] for op in ops]
        high = sorted(b for b in bitpos if b >= low)
        self.groups: List[Any] = []  # int arrays of selected bits, one cross_rdm call each
        chunks = [high[i : i + 7] for i in range(0, len(high), 7)] or [[]]
        where = {}
        for gi, ch in enumerate(chunks):
            self.groups.append((len(ch), _lib.int_array(ch) if ch else None))
            for t, b in enumerate(ch):
                where[b] = (gi * 10 + low + t) * 4
        for b in range(low):
            where[b] = b * 4  # the low bits are in every tile: read them from the first call
        src = []
        for b in bitpos:
            src += [where[b] + c for c in range(4)]
        self.src = torch.tensor(src, dtype=torch.long, device=device)
        dst = []
        for off in dense_offs:
            dst += [off + c for c in range(4)]
        self.dst = torch.tensor(dst, dtype=torch.long, device=device)
        self.m = len(ops)
        structure = [(op.qubits, ("diag",) if op.kind[0] in ("diag", "diagvec") else ("dense",), 4) for op in reversed(ops)]
        self.sub = svengine.compile_circuit(nq, structure, device, absorb_prefix=False)
        sub_idx: List[int] = []
        for off in reversed(dense_offs):
            sub_idx += list(range(off, off + 4))
        self.sub_idx = torch.tensor(sub_idx, dtype=torch.long, device=device)

    def backward(self, nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor, g_all: torch.Tensor) -> None:
        """lam, psi: [2^n] or a contiguous batch [B, 2^n]; dag [.., T], g_all [.., T, 2] likewise."""
        self.unapply(lam, psi, dag)
        self.grads(nbits, lam, psi, dag, g_all)

    def unapply(self, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor) -> None:
        nb = 1 if lam.dim() == 1 else int(lam.shape[0])
        sub_gb = dag.reshape(nb, -1)[:, self.sub_idx].contiguous()
        self.sub.run(psi, sub_gb, batch=nb, gate_batch_stride=int(sub_gb.shape[1]))  # psi_0, lam_0: before the run
        self.sub.run(lam, sub_gb, batch=nb, gate_batch_stride=int(sub_gb.shape[1]))

    def grads(self, nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor, g_all: torch.Tensor) -> None:
        """From the states BEFORE the run."""
        nb = 1 if lam.dim() == 1 else int(lam.shape[0])
        dag2, g2 = dag.reshape(nb, -1), g_all.reshape(nb, -1, 2)
        ng = len(self.groups)
        cr = torch.zeros(nb, ng * 40, 2, dtype=torch.float64, device=psi.device)
        for gi, (nsel, sel) in enumerate(self.groups):
            _lib.call("tcb_sv_cross_rdm", lam.data_ptr(), psi.data_ptr(), nbits, nb, nsel, sel, int(gi > 0),
                      cr.data_ptr() + gi * 40 * 16, ng * 40, _lib.stream_ptr())  # fmt: skip
        c = torch.view_as_complex(cr[:, self.src]).reshape(nb * self.m, 2, 2)
        u = dag2[:, self.dst].reshape(nb * self.m, 2, 2).conj().transpose(1, 2).to(torch.complex128)  # (U^dagger)^dagger
        g = torch.bmm(u, c).reshape(nb, -1)
        g2.index_add_(1, self.dst, torch.view_as_real(g))


class _FusedUnapply:
    """A one-qubit run immediately followed (program order) by a diagonal run, un-applied as ONE sub-circuit: the
    diagonal run's gradients need the states after it, the one-qubit run's the states before it, so nothing in
    between is needed and the two share their passes."""

    def __init__(self, cc: "svengine.CompiledCircuit", first: int, last: int, dense_offs: List[int],
                 device: torch.device) -> None:  # fmt: skip
        nq = cc.plan.nbits
        ops = cc.ops[first:last]
        structure = [(op.qubits, ("diag",) if op.kind[0] in ("diag", "diagvec") else ("dense",), (1 << op.k) ** 2)
                     for op in reversed(ops)]  # fmt: skip
        self.sub = svengine.compile_circuit(nq, structure, device, absorb_prefix=False)
        sub_idx: List[int] = []
        for op, off in zip(reversed(ops), reversed(dense_offs)):
            sub_idx += list(range(off, off + (1 << op.k) ** 2))
        self.sub_idx = torch.tensor(sub_idx, dtype=torch.long, device=device)

    unapply = _OneQubitRun.unapply


class _ConstRun(_FusedUnapply):
    """A run of consecutive gates that take no gradient (H, CNOT, CZ, SWAP ... between the trainable layers): the
    adjoint walk only has to un-apply them, which the fused pass kernel does for the whole run at once."""

    def __init__(self, cc: "svengine.CompiledCircuit", first: int, last: int, dense_offs: List[int],
                 device: torch.device) -> None:  # fmt: skip
        super().__init__(cc, first, last, dense_offs, device)
        self.first, self.last = first, last

    def backward(self, nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor, g_all: torch.Tensor) -> None:
        self.unapply(lam, psi, dag)


class _AdjointTables:
    """Per-circuit index tables for the backward walk, built once and cached on the compiled circuit:
    a dense row-major U^dagger for every gate comes from ONE gather of conj(gate buffer), and the
    float64 per-gate reductions land in one buffer that is scattered back with ONE index_add."""

    def __init__(self, cc: "svengine.CompiledCircuit", nelem: int, device: torch.device, const_mask: Any = None) -> None:
        nq = cc.plan.nbits
        const = tuple(const_mask) if const_mask is not None else (False,) * len(cc.ops)
        dag_idx: List[int] = []
        scat_src: List[int] = []
        scat_dst: List[int] = []
        self.items: List[Any] = []  # (k, bit positions, dense offset) in program order
        for op in cc.ops:
            d = 1 << op.k
            wide = op.k > 2  # constants (toffoli, fredkin ...): un-applied on both states, no gradient block
            off = len(dag_idx)
            for r in range(d):
                for c in range(d):
                    if op.kind[0] == "diagvec":  # U^dagger[r, c] = conj(v[r]) on the diagonal, else the zero sentinel
                        dag_idx.append(op.mat_off + r if r == c else nelem)
                        if r == c and not wide:
                            scat_src.append(off + r * d + c)
                            scat_dst.append(op.mat_off + r)
                    else:  # U^dagger[r, c] = conj(U[c, r])
                        dag_idx.append(op.mat_off + c * d + r)
                        if not wide:
                            scat_src.append(off + r * d + c)
                            scat_dst.append(op.mat_off + r * d + c)
            self.items.append((op.k, _lib.int_array([nq - 1 - q for q in op.qubits]), off))
        self.total = len(dag_idx)
        # program-order segments: ("G", first, last) = gates walked one by one; ("D", first, last) = a run of at
        # least `diag_run_min` consecutive diagonal 1- / 2-qubit gates, differentiated from ONE read of the two
        # states (they commute) and un-applied as one fused sub-circuit.
        self.segments: List[Any] = []
        ops = cc.ops
        offs = [it[2] for it in self.items]

        def add_plain(a: int, b: int) -> None:
            if a >= b:
                return
            if self.segments and isinstance(self.segments[-1], tuple) and self.segments[-1][2] == a:
                self.segments[-1] = ("G", self.segments[-1][1], b)
            else:
                self.segments.append(("G", a, b))

        def add_general(a: int, b: int) -> None:
            """[a, b) holds no long diagonal run: carve out runs of one-qubit gates on distinct qubits and runs of
            constant gates (no gradient wanted: un-applied as one fused sub-circuit, e.g. a CNOT ladder)."""
            i = a
            while i < b:
                j, seen = i, set()
                while j < b and ops[j].k == 1 and ops[j].qubits[0] not in seen:
                    seen.add(ops[j].qubits[0])
                    j += 1
                if j - i >= one_qubit_run_min:
                    self.segments.append(_OneQubitRun(cc, i, j, offs[i:j], device))
                    i = j
                    continue
                # a block of one-qubit gates in which qubits repeat (ry(q) rz(q) per qubit ...): its rounds
                j = i
                while j < b and ops[j].k == 1:
                    j += 1
                if j - i >= 2 * one_qubit_run_min:
                    depth: Dict[int, int] = {}
                    rounds: List[List[int]] = []
                    for g in range(i, j):
                        q = ops[g].qubits[0]
                        r = depth.get(q, 0)
                        depth[q] = r + 1
                        if r == len(rounds):
                            rounds.append([])
                        rounds[r].append(g)
                    if max(len(r) for r in rounds) >= one_qubit_run_min:
                        for r in rounds:
                            if len(r) >= one_qubit_run_min:
                                self.segments.append(_OneQubitRun(cc, 0, 0, [offs[g] for g in r], device, indices=r))
                            else:
                                for g in r:
                                    self.segments.append(("G", g, g + 1))
                        i = j
                        continue
                j = i
                while j < b and const[j] and ops[j].k <= 4:
                    j += 1
                if j - i >= const_run_min:
                    self.segments.append(_ConstRun(cc, i, j, offs[i:j], device))
                    i = j
                else:
                    add_plain(i, i + 1)
                    i += 1

        i = start = 0
        while i < len(ops):
            j = i
            while j < len(ops) and ops[j].kind[0] in ("diag", "diagvec") and ops[j].k <= 2:
                j += 1
            if j - i >= diag_run_min:
                add_general(start, i)
                self.segments.append(_DiagRun(cc, i, j, offs[i:j], device))
                i = start = j
            else:
                i = max(j, i + 1)
        add_general(start, len(ops))
        self.fused: Dict[int, _FusedUnapply] = {}  # index of a _OneQubitRun segment directly followed by a _DiagRun
        for k in range(len(self.segments) - 1):
            a, b = self.segments[k], self.segments[k + 1]
            if isinstance(a, _OneQubitRun) and isinstance(b, _DiagRun) and a.last == b.first:
                self.fused[k] = _FusedUnapply(cc, a.first, b.last, offs[a.first : b.last], device)
        self.dag_idx = torch.tensor(dag_idx, dtype=torch.long, device=device)
        self.scat_src = torch.tensor(scat_src, dtype=torch.long, device=device)
        self.scat_dst = torch.tensor(scat_dst, dtype=torch.long, device=device)


def _adjoint_tables(cc: "svengine.CompiledCircuit", gatebuf: torch.Tensor, const_mask: Any = None) -> _AdjointTables:
    """Cached per circuit, device and set of constant gates (`const_mask[g]`: gate g takes no gradient)."""
    cache = getattr(cc, "_adjoint_tables", None)
    if cache is None:
        cache = cc._adjoint_tables = {}
    key = (str(gatebuf.device), const_mask)
    t = cache.get(key)
    if t is None:
        if len(cache) > 8:
            cache.clear()
        t = cache[key] = _AdjointTables(cc, gatebuf.numel(), gatebuf.device, const_mask)
    return t


class _Evolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, gatebuf: torch.Tensor, init: Optional[torch.Tensor], cc: Any, const_mask: Any = None) -> torch.Tensor:
        state = _forward(cc, gatebuf.detach(), init)
        ctx.cc = cc
        ctx.const_mask = const_mask
        ctx.has_init = init is not None
        ctx.save_for_backward(gatebuf.detach(), state)
        return state

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore[override]
        gatebuf, psi_out = ctx.saved_tensors
        cc = ctx.cc
        nbits = cc.plan.nbits
        if not assume_unitary:
            raise _lib.EngineError("autograd.assume_unitary=False (recompute mode) is not implemented in this round")
        tabs = _adjoint_tables(cc, gatebuf, ctx.const_mask)
        lam = grad_out.to(torch.complex64).resolve_conj().reshape(-1).clone()
        psi = psi_out.clone()
        src = torch.cat([gatebuf.conj().resolve_conj(), torch.zeros(1, dtype=gatebuf.dtype, device=gatebuf.device)])
        dag = src[tabs.dag_idx].contiguous()  # every U^dagger, dense row-major, in program order
        g_all = torch.zeros(max(tabs.total, 1), 2, dtype=torch.float64, device=gatebuf.device)
        # psi_in = U^dagger psi_out; dL/dU += lam_out (x) conj(psi_in); lam_in = U^dagger lam_out — one pass per
        # gate over both states, the whole walk one call (tcb_sv_plan_vjp)
        _walk(cc, tabs, nbits, lam, psi, dag, g_all)
        grad_buf = torch.zeros_like(gatebuf)
        torch.view_as_real(grad_buf).index_add_(0, tabs.scat_dst, g_all[tabs.scat_src].to(torch.float32))
        grad_init = lam if ctx.has_init and ctx.needs_input_grad[1] else None
        return grad_buf, grad_init, None, None


# ---- batched evolution under torch.vmap ------------------------------------------------------------
# `backend.vmap / vvag` run the user's function ONCE under torch.vmap: every torch op on the parameters is
# batched by functorch, and at the engine boundary the batched gate buffer is unwrapped to its physical
# [B, elems] tensor, the kernels run with batch = B (one circuit build, one plan, B states side by side in
# HBM), and the result is wrapped again.  Plain autograd runs on the physical tensors.
_ft = torch._C._functorch


def is_batched(t: Any) -> bool:
    return isinstance(t, torch.Tensor) and _ft.is_batchedtensor(t)


def wants_grad(t: torch.Tensor) -> bool:
    """requires_grad of the physical tensor (the flag is not mirrored on torch.vmap's batched wrappers)."""
    while is_batched(t):
        t = _ft.get_unwrapped(t)
    return bool(t.requires_grad)


def unwrap_batched(t: torch.Tensor) -> Any:
    """(physical tensor with the batch dimension first, vmap level)."""
    lvl, bd = _ft.maybe_get_level(t), _ft.maybe_get_bdim(t)
    phys = _ft.get_unwrapped(t)
    if is_batched(phys):
        raise _lib.EngineError("nested vmap is not supported by the batched engine path")
    return phys.movedim(bd, 0), lvl


def outside_vmap() -> Any:
    from torch._functorch import pyfunctorch

    return pyfunctorch.temporarily_pop_interpreter_stack()


def rewrap_batched(t: torch.Tensor, lvl: int) -> torch.Tensor:
    return _ft._add_batch_dim(t, 0, lvl)


class _EvolveBatched(torch.autograd.Function):
    """B independent parameter sets of one circuit: states [B, 2^n], gate buffers [B, elems]."""

    @staticmethod
    def forward(ctx: Any, gatebuf: torch.Tensor, cc: Any, const_mask: Any = None) -> torch.Tensor:
        ctx.const_mask = const_mask
        gb = gatebuf.detach().to(torch.complex64).contiguous()
        nb, nelem = gb.shape
        nbits = cc.plan.nbits
        if gb.is_cuda:
            free, _ = torch.cuda.mem_get_info(gb.device)
            need = 5 * (nb << nbits) * 8  # state, saved copy, lam, psi, H psi / cotangents
            if need > 0.9 * free:
                raise _lib.EngineError(f"a batch of {nb} {nbits}-qubit states does not fit in free HBM ({free >> 30} GiB)")
        state = torch.empty(nb << nbits, dtype=torch.complex64, device=gb.device)
        _lib.require_cuda(state, "state")
        _lib.call("tcb_sv_init_zero", state.data_ptr(), nbits, nb, _lib.stream_ptr())
        cc.run(state, gb, batch=nb, gate_batch_stride=nelem)
        state = state.reshape(nb, 1 << nbits)
        ctx.cc = cc
        ctx.save_for_backward(gb, state)
        return state

    @staticmethod
    def backward(ctx: Any, grad_out: torch.Tensor):  # type: ignore[override]
        gb, psi_out = ctx.saved_tensors
        cc = ctx.cc
        nbits = cc.plan.nbits
        nb, nelem = gb.shape
        tabs = _adjoint_tables(cc, gb[0], ctx.const_mask)
        lam = grad_out.to(torch.complex64).resolve_conj().contiguous().clone()
        psi = psi_out.clone()
        src = torch.cat([gb.conj().resolve_conj(), torch.zeros(nb, 1, dtype=gb.dtype, device=gb.device)], dim=1)
        dag = src[:, tabs.dag_idx].contiguous()
        g_all = torch.zeros(nb, max(tabs.total, 1), 2, dtype=torch.float64, device=gb.device)
        _walk(cc, tabs, nbits, lam, psi, dag, g_all)
        grad_buf = torch.zeros(nb, nelem, 2, dtype=torch.float32, device=gb.device)
        grad_buf.index_add_(1, tabs.scat_dst, g_all[:, tabs.scat_src].to(torch.float32))
        return torch.view_as_complex(grad_buf), None, None


def _walk(cc: Any, tabs: "_AdjointTables", nbits: int, lam: torch.Tensor, psi: torch.Tensor, dag: torch.Tensor,
          g_all: torch.Tensor) -> None:  # fmt: skip
    """The adjoint walk (one sample [2^n], or a contiguous batch [B, 2^n] walked together): psi is un-computed in
    place, lam becomes the cotangent of the initial state, g_all (float64 pairs, one dense block per gate)
    receives dL/dU."""

    def plain(first: int, last: int) -> None:
        if lam.dim() == 1:
            cc.vjp(lam, psi, dag, g_all, first, last)
        else:
            for b in range(lam.shape[0]):
                cc.vjp(lam[b], psi[b], dag[b], g_all[b], first, last)

    if layered_adjoint and len(tabs.segments) > 1:
        k = len(tabs.segments) - 1
        while k >= 0:
            seg = tabs.segments[k]
            if isinstance(seg, _DiagRun) and (k - 1) in tabs.fused:
                seg.grads(nbits, lam, psi, dag, g_all)             # states after the diagonal run
                tabs.fused[k - 1].unapply(lam, psi, dag)           # both runs in one fused sub-circuit
                tabs.segments[k - 1].grads(nbits, lam, psi, dag, g_all)  # states before the one-qubit run
                k -= 2
                continue
            if isinstance(seg, (_DiagRun, _OneQubitRun, _ConstRun)):
                seg.backward(nbits, lam, psi, dag, g_all)
            else:
                plain(seg[1], seg[2])
            k -= 1
    else:
        plain(0, len(cc.ops))


def evolve(cc: "svengine.CompiledCircuit", gatebuf: torch.Tensor, init: Optional[torch.Tensor],
           const_mask: Any = None) -> torch.Tensor:  # fmt: skip
    """`const_mask[g]` (optional, one bool per gate of cc.ops): gate g is a constant — the backward walk skips its
    gradient and un-applies runs of such gates as fused sub-circuits."""
    if is_batched(gatebuf) or is_batched(init):
        if init is not None or cc.prefix_levels:
            raise _lib.EngineError("the batched engine path starts from |0...0> (no `inputs`, no absorbed prefix)")
        phys, lvl = unwrap_batched(gatebuf)
        with outside_vmap():
            out = _EvolveBatched.apply(phys, cc, const_mask)
        return rewrap_batched(out, lvl)
    needs = torch.is_grad_enabled() and (gatebuf.requires_grad or (init is not None and init.requires_grad))
    if not needs:
        return _forward(cc, gatebuf, init)
    return _Evolve.apply(gatebuf, init, cc, const_mask)
