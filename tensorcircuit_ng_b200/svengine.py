"""
Statevector engine: recognises a circuit-shaped tensor network at the contractor
boundary, compiles it into fused HBM passes (passplan.py) and runs them through the
C ABI (`tcb_sv_*`, include/tcb200.h).

What it replaces in the reference: the `for (a, b) in path: tn.contract_between(...)`
loop plus the final `reorder_edges` of tensorcircuit/cons.py:937-960 when the network
handed to the contractor is `Circuit._copy()` (tensorcircuit/circuit.py:711-712) —
i.e. `wavefunction()` / the cached state of `expectation*()`.
"""

from __future__ import annotations

import math
import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, passplan
from .passplan import GateOp, GlobalStep, PassStep, Plan

_MAX_STATE_QUBITS = 34  # 2^34 complex64 = 128 GiB: upper bound for one B200


class NotCircuitShaped(Exception):
    """The network is not `inputs -> gates -> dangling outputs`; use the TN path."""


class NonUnitaryGradient(NotCircuitShaped):
    """A gradient is wanted through a gate that is not unitary: the adjoint-method backward (which un-computes
    the state with U^dagger) does not apply, the tensor-network route with ordinary autograd does."""


unitarity_tol = 1e-4


def _require_unitary_for_adjoint(gates: Sequence[Tuple[Any, Tuple[int, ...], bool]]) -> None:
    """One batched device check  max |U^dagger U - 1|  over the gates whose factory does not guarantee unitarity
    (`any`, `diagonal`, `exp` / `exp1` with a tensor generator or complex time, foreign nodes)."""
    from . import autograd

    worst = None
    for node, qubits, packed in gates:
        if getattr(node, "_b200_unitary", False):
            continue
        t = node.tensor
        while autograd.is_batched(t):
            t = torch._C._functorch.get_unwrapped(t)
        t = t.detach().to(torch.complex64)
        d = 1 << len(qubits)
        if packed:
            dev = (t.reshape(-1, d).abs() ** 2 - 1.0).abs().max()
        else:
            m = t.reshape(-1, d, d)
            dev = (m.conj().transpose(1, 2) @ m - torch.eye(d, dtype=m.dtype, device=m.device)).abs().max()
        worst = dev if worst is None else torch.maximum(worst, dev.to(worst.device))
    if worst is not None and float(worst) > unitarity_tol:
        raise NonUnitaryGradient(f"non-unitary gate (max |U^dagger U - 1| = {float(worst):.2e}) under autograd")


# ---------------------------------------------------------------------------------------
def _is_copynode(node: Any) -> bool:
    return type(node).__name__ == "CopyNode"


def _is_input_node(node: Any) -> bool:
    flag = getattr(node, "flag", "")
    if isinstance(flag, str) and flag.startswith("inputs"):
        return True
    return node.get_rank() == 1 and str(getattr(node, "name", "")).startswith("qb-")


def _other_end(edge: Any, node: Any, axis: int) -> Tuple[Any, int]:
    if edge.node1 is node and edge.axis1 == axis:
        return edge.node2, edge.axis2
    return edge.node1, edge.axis1


def extract_gate_stream(nodes: Sequence[Any], output_edge_order: Sequence[Any],
                        max_qubits: int = _MAX_STATE_QUBITS):
    """Walk every output wire back to its input; returns (n, init_node or None, gate list).

    gate list entries: (node, qubits tuple, packed_diagonal flag) sorted by creation order
    (`_stable_id_`, tensorcircuit/cons.py:28-53 — the only record of program order the
    contractor sees).  Raises NotCircuitShaped on anything else (expectation sandwiches,
    amplitude networks, MPS inputs, ...).
    """
    n = len(output_edge_order)
    if n == 0 or n > max_qubits:
        raise NotCircuitShaped("no dangling outputs" if n == 0 else "too many qubits for a statevector")
    node_ids = {id(x) for x in nodes}
    # classify every node once (the wire walk below visits a gate once per leg)
    cls: Dict[int, Tuple[bool, bool, int]] = {id(x): (_is_copynode(x), _is_input_node(x), x.get_rank()) for x in nodes}
    legs: Dict[int, List[Optional[int]]] = {}
    gate_nodes: Dict[int, Any] = {}
    diag_legs: Dict[int, List[Optional[int]]] = {}
    init_node = None
    n_zero_inputs = 0
    visited_inputs = set()
    for q, e in enumerate(output_edge_order):
        if not e.is_dangling():
            raise NotCircuitShaped("output edge is not dangling")
        node, axis = e.node1, e.axis1
        steps = 0
        while True:
            steps += 1
            if steps > 1_000_000:
                raise NotCircuitShaped("wire walk did not terminate")
            if id(node) not in node_ids:
                raise NotCircuitShaped("wire leaves the node set")
            is_copy, is_input, r = cls[id(node)]
            if is_copy:
                # diagonal gate in hyperedge form (tensorcircuit/basecircuit.py:343-355):
                # cn[0] <- previous front, cn[1] <-> coefficient leg, cn[2] -> next
                if r != 3 or axis != 2:
                    raise NotCircuitShaped("unsupported CopyNode wiring")
                coef, cax = _other_end(node.edges[1], node, 1)
                if coef is None or _is_copynode(coef):
                    raise NotCircuitShaped("CopyNode without coefficient node")
                dl = diag_legs.setdefault(id(coef), [None] * coef.get_rank())
                gate_nodes[id(coef)] = coef
                if dl[cax] is not None:
                    raise NotCircuitShaped("coefficient leg used twice")
                dl[cax] = q
                prev = node.edges[0]
                if prev.is_dangling():
                    raise NotCircuitShaped("dangling CopyNode input")
                node, axis = _other_end(prev, node, 0)
                continue
            if is_input:
                if r == 1 and str(getattr(node, "name", "")).startswith("qb-"):
                    n_zero_inputs += 1  # |0> leaf of all_zero_nodes (basecircuit.py:52-66)
                else:
                    if init_node is not None and init_node is not node:
                        raise NotCircuitShaped("several multi-leg input nodes")
                    if axis != q:
                        raise NotCircuitShaped("input legs are permuted")
                    init_node = node
                if (id(node), axis) in visited_inputs:
                    raise NotCircuitShaped("input leg reached twice")
                visited_inputs.add((id(node), axis))
                break
            if r % 2:
                raise NotCircuitShaped("odd-rank node on a wire")
            k = r // 2
            if axis >= k:
                raise NotCircuitShaped("wire enters a gate through an input leg")
            lg = legs.setdefault(id(node), [None] * k)
            gate_nodes[id(node)] = node
            if lg[axis] is not None:
                raise NotCircuitShaped("gate output leg used twice")
            lg[axis] = q
            ine = node.edges[axis + k]
            if ine.is_dangling():
                raise NotCircuitShaped("dangling gate input")
            node, axis = _other_end(ine, node, axis + k)
    if init_node is not None and n_zero_inputs:
        raise NotCircuitShaped("mixed input kinds")
    if init_node is not None and init_node.get_rank() != n:
        raise NotCircuitShaped("input node rank mismatch")
    # every node must have been accounted for
    seen = set(legs) | set(diag_legs)
    gates: List[Tuple[Any, Tuple[int, ...], bool]] = []
    for x in nodes:
        if cls[id(x)][0] or cls[id(x)][1]:
            continue
        if id(x) not in seen:
            raise NotCircuitShaped("node not on any wire")
        if id(x) in legs:
            lg = legs[id(x)]
            if any(v is None for v in lg):
                raise NotCircuitShaped("gate with unvisited legs")
            gates.append((x, tuple(int(v) for v in lg), False))  # type: ignore[arg-type]
        else:
            dl = diag_legs[id(x)]
            if any(v is None for v in dl):
                raise NotCircuitShaped("diagonal gate with unvisited legs")
            gates.append((x, tuple(int(v) for v in dl), True))  # type: ignore[arg-type]
    gates.sort(key=lambda g: getattr(g[0], "_stable_id_", -1))
    return n, init_node, gates


# ---- gate classification at a nodes-only boundary (SURVEY §7 "Diagonal / controlled detection") ---------
# Nodes built by this package's gates.py carry `_b200_kind`.  Nodes coming from a real TensorCircuit-NG install
# carry only `name` (overwritten with the method name, tensorcircuit/basecircuit.py:278 — user-overridable) and
# their tensor.  For those:
#   * a name of a gate family that is dense for generic parameters (rx, ry, h, u, r, swap ...) gives "dense"
#     without looking at values — the safe direction, a dense treatment is correct for any matrix;
#   * everything else (rz / cz / cnot / exp1 / any / unknown names ...) is classified by a STRUCTURAL probe of the
#     tensor's exact-zero pattern: diagonal, or a 2x2 block controlled by the first k-1 qubits, else dense.  The
#     probe reads the foreign tensors back ONCE per network topology (one batched device->host copy) and its result
#     is cached with that topology, so a training loop pays no synchronisation after its first step.  Zero patterns
#     are parameter independent except on measure-zero parameter values; an exactly-identity matrix is treated as
#     dense, so a parametrised gate that happens to start at theta = 0 is not mistaken for a diagonal one.
#   * `trust_gate_names = True` skips the probe for the reference's own names below.
_NAME_KINDS: Dict[str, Tuple[Any, ...]] = {
    "z": ("diag",), "s": ("diag",), "t": ("diag",), "sd": ("diag",), "td": ("diag",), "i": ("diag",),
    "rz": ("diag",), "phase": ("diag",), "cz": ("diag",), "rzz": ("diag",), "cphase": ("diag",),
    "crz": ("diag",), "oz": ("diag",), "orz": ("diag",),
    "cnot": ("ctrl", 1, 1), "cx": ("ctrl", 1, 1), "cy": ("ctrl", 1, 1), "crx": ("ctrl", 1, 1),
    "cry": ("ctrl", 1, 1), "cu": ("ctrl", 1, 1), "cr": ("ctrl", 1, 1), "toffoli": ("ctrl", 2, 3),
    "ox": ("ctrl", 1, 0), "oy": ("ctrl", 1, 0), "orx": ("ctrl", 1, 0), "ory": ("ctrl", 1, 0),
}  # fmt: skip
_DENSE_NAMES = frozenset(["x", "y", "h", "rx", "ry", "r", "u", "wroot", "swap", "iswap", "rxx", "ryy", "fredkin"])

trust_gate_names = False  # opt-in: node names are user-overridable in the reference
_probe_cache: Dict[Any, Dict[int, Tuple[Any, ...]]] = {}
probe_readbacks = 0  # device->host probe copies issued so far (tests pin "once per topology")


def classify_matrix(m: np.ndarray) -> Tuple[Any, ...]:
    """Structural kind of a 2^k x 2^k matrix from its exact-zero pattern."""
    d = m.shape[0]
    k = int(round(math.log2(d)))
    off = m - np.diag(np.diagonal(m))
    if not off.any():
        return ("dense",) if np.all(np.diagonal(m) == 1) else ("diag",)
    if 2 <= k <= 3:
        nb = d // 2
        blocks = [m[2 * b : 2 * b + 2, 2 * b : 2 * b + 2] for b in range(nb)]
        rest = m.copy()
        for b in range(nb):
            rest[2 * b : 2 * b + 2, 2 * b : 2 * b + 2] = 0
        if not rest.any():
            eye = np.eye(2, dtype=m.dtype)
            active = [b for b in range(nb) if not np.array_equal(blocks[b], eye)]
            if len(active) == 1:
                nctrl, b = k - 1, active[0]
                pol = 0
                for ci in range(nctrl):  # bit ci of pol = required value of control ci (control 0 = block-index MSB)
                    pol |= ((b >> (nctrl - 1 - ci)) & 1) << ci
                return ("ctrl", nctrl, pol)
    return ("dense",)


def gate_kind(node: Any, packed_diag: bool) -> Tuple[Any, ...]:
    """Kind of one node WITHOUT looking at values (None: needs the structural probe)."""
    if packed_diag:
        return ("diagvec",)
    kind = getattr(node, "_b200_kind", None)
    if kind is not None:
        return tuple(kind) if kind[0] != "diagvec" else ("dense",)
    name = str(getattr(node, "name", ""))
    if name in _DENSE_NAMES:
        return ("dense",)
    if trust_gate_names and name in _NAME_KINDS:
        return _NAME_KINDS[name]
    return None  # type: ignore[return-value]


def gate_kinds(gates: Sequence[Tuple[Any, Tuple[int, ...], bool]]) -> List[Tuple[Any, ...]]:
    """Kinds of a whole gate stream; foreign nodes are probed once per topology (see above)."""
    global probe_readbacks
    kinds: List[Any] = [gate_kind(g[0], g[2]) for g in gates]
    todo = [i for i, k in enumerate(kinds) if k is None]
    if not todo:
        return kinds
    key = tuple((g[1], str(getattr(g[0], "name", "")), tuple(g[0].shape), g[2]) for g in gates)
    hit = _probe_cache.get(key)
    if hit is None:
        from . import autograd

        flat = []
        for i in todo:
            t = gates[i][0].tensor
            while autograd.is_batched(t):  # under vmap: the structure of the first sample stands for the batch
                t = torch._C._functorch.get_unwrapped(t)
            t = t.detach()
            d = 1 << len(gates[i][1])
            flat.append(t.reshape(-1, d * d)[0].to(torch.complex64))
        host = torch.cat(flat).cpu().numpy()  # ONE device->host copy for all foreign gates
        probe_readbacks += 1
        hit, off = {}, 0
        for i in todo:
            d = 1 << len(gates[i][1])
            hit[i] = classify_matrix(host[off : off + d * d].reshape(d, d))
            off += d * d
        if len(_probe_cache) > 256:
            _probe_cache.clear()
        _probe_cache[key] = hit
    for i in todo:
        kinds[i] = hit[i]
    return kinds


# ---------------------------------------------------------------------------------------
def split_prefix(ops: Sequence[GateOp], nq: int) -> Tuple[List[List[GateOp]], List[GateOp]]:
    """(per-qubit leading 1q gates, remaining gates).  The 1q gates a qubit sees before its first
    multi-qubit gate act on |0> alone: they fold into a per-qubit 2-vector (what the reference's
    `_merge_single_gates` does to its input nodes, tensorcircuit/cons.py:298-374)."""
    open_ = [True] * nq
    prefix: List[List[GateOp]] = [[] for _ in range(nq)]
    rest: List[GateOp] = []
    for g in ops:
        if g.k == 1 and open_[g.qubits[0]] and g.kind[0] in ("dense", "diag", "diagvec"):
            prefix[g.qubits[0]].append(g)
        else:
            for q in g.qubits:
                open_[q] = False
            rest.append(g)
    return prefix, rest


class CompiledCircuit:
    """A plan plus its device-resident programs (uploaded once, reused every step)."""

    def __init__(self, plan: Plan, ops: List[GateOp], device: torch.device,
                 prefix: Optional[List[List[GateOp]]] = None, nq: Optional[int] = None) -> None:  # fmt: skip
        self.plan = plan
        self.ops = ops  # the WHOLE circuit in program order (the adjoint backward pass walks it)
        self.device = device
        # product-state start: per level, (flat bit positions, [m, 4] indices into cat(gatebuf, 0))
        self.prefix_levels: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.nq = nq if nq is not None else plan.nbits
        if prefix is not None and any(prefix):
            depth = max(len(p) for p in prefix)
            for lvl in range(depth):
                pos, idx = [], []
                for q, gl in enumerate(prefix):
                    if lvl < len(gl):
                        g = gl[lvl]
                        pos.append(self.nq - 1 - q)
                        if g.kind[0] == "diagvec":
                            idx.append([g.mat_off, -1, -1, g.mat_off + 1])
                        else:
                            idx.append([g.mat_off, g.mat_off + 1, g.mat_off + 2, g.mat_off + 3])
                self.prefix_levels.append((torch.tensor(pos, dtype=torch.long, device=device),
                                           torch.tensor(idx, dtype=torch.long, device=device)))  # fmt: skip
        chunks = [s.program for s in plan.steps if isinstance(s, PassStep)]
        self.offsets: List[int] = []
        off = 0
        for c in chunks:
            self.offsets.append(off)
            off += len(c)
        if chunks:
            host = np.concatenate(chunks).astype(np.int32)
            self.programs = torch.from_numpy(host).to(device)
        else:
            host = np.zeros(0, dtype=np.int32)
            self.programs = torch.zeros(1, dtype=torch.int32, device=device)
        self._programs_host = host
        self._handle: Optional[ctypes.c_void_p] = None
        self.has_vjp = all(g.k <= 7 for g in ops)

    # -- the native plan object (include/tcb200.h: tcb_sv_plan_*) ------------------------------------
    def handle(self) -> ctypes.c_void_p:
        """Created on first use: the library copies the pass programs to the device and keeps the step /
        gate tables, so a circuit execution (or its whole adjoint walk) is ONE call through the C ABI."""
        if self._handle is not None:
            return self._handle
        rows: List[List[int]] = []
        pi = 0
        for step in self.plan.steps:
            if isinstance(step, PassStep):
                rows.append([0, self.offsets[pi], len(step.program), step.tile_bits, step.low_bits, step.pool_elems,
                             0, 0, 0] + [0] * 7)  # fmt: skip
                pi += 1
            else:
                g = step.gate
                bp = list(step.bitpos) + [0] * (7 - len(step.bitpos))
                if g.is_diag:
                    stride = 1 if g.kind[0] == "diagvec" else (1 << g.k) + 1
                    rows.append([2, 0, 0, 0, 0, 0, g.k, g.mat_off, stride] + bp)
                else:
                    rows.append([1, 0, 0, 0, 0, 0, g.k, g.mat_off, 0] + bp)
        steps = np.asarray(rows, dtype=np.int64).reshape(-1, 16) if rows else np.zeros((0, 16), np.int64)
        grows, off = [], 0
        if self.has_vjp:
            nb = self.plan.nbits
            for g in self.ops:
                bp = [nb - 1 - q for q in g.qubits]
                grows.append([g.k, off] + bp + [0] * (8 - len(bp)))
                off += (1 << g.k) ** 2
        gates = np.asarray(grows, dtype=np.int64).reshape(-1, 10) if grows else np.zeros((0, 10), np.int64)
        h = ctypes.c_void_p()
        host = np.ascontiguousarray(self._programs_host)
        _lib.check(_lib.load().tcb_sv_plan_create(
            self.plan.nbits, host.ctypes.data if len(host) else None, len(host),
            steps.ctypes.data if len(steps) else None, len(steps),
            gates.ctypes.data if len(gates) else None, len(gates), ctypes.byref(h)))  # fmt: skip
        self._handle = h
        self._n_launch = int(_lib.load().tcb_sv_plan_launches(h, 0))
        self._n_launch_vjp = int(_lib.load().tcb_sv_plan_launches(h, 1))
        return h

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().tcb_sv_plan_destroy(h)
            except Exception:  # pylint: disable=broad-except  (interpreter shutdown)
                pass
            self._handle = None

    def vjp(self, lam: torch.Tensor, psi: torch.Tensor, udag: torch.Tensor, grad: torch.Tensor,
            first: int = 0, last: Optional[int] = None) -> None:  # fmt: skip
        """The whole adjoint walk (last gate first) in one call: psi is un-computed in place, lam becomes the
        cotangent of the initial state, grad (float64 pairs, dense block per gate) accumulates dL/dU."""
        if not self.has_vjp:
            k = max(g.k for g in self.ops)
            raise _lib.EngineError(f"the adjoint walk supports gates of up to 7 qubits (found {k})")
        h = self.handle()
        last = len(self.ops) if last is None else last
        _lib.check(_lib.load().tcb_sv_plan_vjp_range(h, first, last, lam.data_ptr(), psi.data_ptr(), udag.data_ptr(),
                                                     grad.data_ptr(), _lib.stream_ptr()))  # fmt: skip
        _lib.launch_count += sum(1 if g.k <= 2 else 2 for g in self.ops[first:last])

    def _product_vectors(self, gatebuf: torch.Tensor) -> torch.Tensor:
        """[nq, 2] complex64 by flat bit position: what the absorbed leading 1q gates make of |0> on every qubit."""
        v = torch.zeros(self.nq, 2, dtype=torch.complex64, device=self.device)
        v[:, 0] = 1.0
        if self.prefix_levels:
            gb = torch.cat([gatebuf.detach(), torch.zeros(1, dtype=gatebuf.dtype, device=gatebuf.device)])
            for pos, idx in self.prefix_levels:
                m = gb[idx].reshape(-1, 2, 2)
                v[pos] = torch.bmm(m, v[pos].unsqueeze(-1)).squeeze(-1)
        return v

    def start(self, state: torch.Tensor, gatebuf: torch.Tensor) -> None:
        """Write the initial state of the compiled part: |0...0>, or the product state that the
        absorbed leading 1q gates make of it (one write pass, no read)."""
        nbits = self.plan.nbits
        if not self.prefix_levels:
            _lib.call("tcb_sv_init_zero", state.data_ptr(), nbits, 1, _lib.stream_ptr())
            return
        v = self._product_vectors(gatebuf)
        _lib.call("tcb_sv_init_product", state.data_ptr(), nbits, v.data_ptr(), self.nq, 0, _lib.stream_ptr())

    def start_and_run(self, state: torch.Tensor, gatebuf: torch.Tensor) -> None:
        """`start` + `run` for one state.  When the plan opens with a fused pass of the default tile size, that pass
        GENERATES its tiles from the product vectors (`tcb_sv_run_pass_generate`): the initial state is never written
        to or read from HBM (one write pass and half a pass of traffic less per evolution)."""
        steps = self.plan.steps
        if not (fuse_start and steps and isinstance(steps[0], PassStep) and steps[0].tile_bits == 12):
            self.start(state, gatebuf)
            self.run(state, gatebuf)
            return
        _lib.require_cuda(state, "state")
        _lib.require_cuda(gatebuf, "gate buffer")
        v = self._product_vectors(gatebuf)
        st0 = steps[0]
        _lib.call("tcb_sv_run_pass_generate", state.data_ptr(), self.plan.nbits, self.programs.data_ptr(),
                  len(st0.program), st0.tile_bits, st0.low_bits, st0.pool_elems, gatebuf.data_ptr(), 0, v.data_ptr(),
                  self.nq, _lib.stream_ptr())  # fmt: skip
        self.run_steps(state, gatebuf, first=1)

    def run(self, state: torch.Tensor, gatebuf: torch.Tensor, batch: int = 1, gate_batch_stride: int = 0,
            index_base: int = 0) -> None:  # fmt: skip
        """Apply the compiled part of the circuit IN PLACE to `state` ([batch * 2^nbits] complex64 on
        the GPU).  With absorbed leading gates the state must come from `start`."""
        _lib.require_cuda(state, "state")
        _lib.require_cuda(gatebuf, "gate buffer")
        if use_native_plans:
            h = self.handle()
            _lib.check(_lib.load().tcb_sv_plan_execute(h, state.data_ptr(), batch, gatebuf.data_ptr(),
                                                       gate_batch_stride, index_base, _lib.stream_ptr()))  # fmt: skip
            _lib.launch_count += self._n_launch
            return
        # step by step through the kernel-level entry points (TCB_NATIVE_PLANS=0; same launches)
        self.run_steps(state, gatebuf, batch, gate_batch_stride, index_base)

    def run_steps(self, state: torch.Tensor, gatebuf: torch.Tensor, batch: int = 1, gate_batch_stride: int = 0,
                  index_base: int = 0, first: int = 0) -> None:  # fmt: skip
        """Steps [first:] of the plan, one C-ABI call per step, in place."""
        nbits = self.plan.nbits
        stream = _lib.stream_ptr()
        sp = state.data_ptr()
        gp = gatebuf.data_ptr()
        pi = 0
        for si, step in enumerate(self.plan.steps):
            if isinstance(step, PassStep):
                prog_ptr = self.programs.data_ptr() + 4 * self.offsets[pi]
                pi += 1
                if si >= first:
                    _lib.call("tcb_sv_run_pass", sp, nbits, batch, prog_ptr, len(step.program), step.tile_bits,
                              step.low_bits, step.pool_elems, gp, gate_batch_stride, index_base, stream)  # fmt: skip
            elif si >= first:
                g = step.gate
                bp = _lib.int_array(step.bitpos)
                mp = gp + 8 * g.mat_off
                if g.is_diag:
                    stride = 1 if g.kind[0] == "diagvec" else (1 << g.k) + 1
                    _lib.call("tcb_sv_apply_diag", sp, nbits, batch, bp, g.k, mp, stride, gate_batch_stride,
                              index_base, stream)  # fmt: skip
                else:
                    _lib.call("tcb_sv_apply_dense", sp, nbits, batch, bp, g.k, mp, gate_batch_stride, stream)


_plan_cache: Dict[Any, CompiledCircuit] = {}
import os as _os

use_native_plans = _os.environ.get("TCB_NATIVE_PLANS", "1") != "0"
fuse_start = _os.environ.get("TCB_FUSE_START", "1") != "0"  # 0: always init kernel + passes (CompiledCircuit.start_and_run)

auto_low_bits = "TCB_LOW_BITS" not in _os.environ  # unset: the planner may pick L = 2 when it saves a pass
plan_options: Dict[str, Any] = {
    "tile_bits": int(_os.environ.get("TCB_TILE_BITS", "12")),
    "low_bits": int(_os.environ.get("TCB_LOW_BITS", "3")),
}


def compile_circuit(nq: int, structure: Sequence[Tuple[Tuple[int, ...], Tuple[Any, ...], int]],
                    device: torch.device, nbits_local: Optional[int] = None,
                    absorb_prefix: bool = False) -> CompiledCircuit:  # fmt: skip
    """structure: per gate (qubits, kind, numel of its tensor). Cached by structure.
    `absorb_prefix`: the circuit starts from |0...0>, fold each qubit's leading 1q gates into the
    initial product state (CompiledCircuit.start)."""
    key = (nq, nbits_local, tuple(structure), str(device), tuple(sorted(plan_options.items())), absorb_prefix)
    cc = _plan_cache.get(key)
    if cc is None:
        ops: List[GateOp] = []
        off = 0
        for gi, (qubits, kind, numel) in enumerate(structure):
            ops.append(GateOp(tuple(qubits), tuple(kind), off, gi))
            off += numel
        prefix = None
        plan_ops: List[GateOp] = ops
        if absorb_prefix and nbits_local is None:
            prefix, plan_ops = split_prefix(ops, nq)
        plan = passplan.compile_plan(plan_ops, nq, nbits_local=nbits_local, **plan_options)
        if auto_low_bits and nbits_local is None and plan_options.get("low_bits") == 3 and nq >= 28:  # (measured there)
            # 16-byte segments (L = 2) leave one more free tile bit per pass: a pass is ~3 % slower but long
            # circuits sometimes need one pass fewer (30-qubit QAOA p=8: 18 instead of 19, 151 vs 155 ms)
            n3 = sum(isinstance(st, PassStep) for st in plan.steps)
            if n3 >= 4:
                alt = passplan.compile_plan(plan_ops, nq, nbits_local=nbits_local, **{**plan_options, "low_bits": 2})
                n2 = sum(isinstance(st, PassStep) for st in alt.steps)
                if len(alt.steps) - n2 <= len(plan.steps) - n3 and n2 * 1.04 < n3:
                    plan = alt
        cc = CompiledCircuit(plan, ops, device, prefix=prefix, nq=nq)
        if len(_plan_cache) > 256:
            _plan_cache.clear()
        _plan_cache[key] = cc
    return cc


def new_zero_state(nbits: int, batch: int, device: torch.device) -> torch.Tensor:
    state = torch.empty(batch << nbits, dtype=torch.complex64, device=device)
    _lib.require_cuda(state, "state")
    _lib.call("tcb_sv_init_zero", state.data_ptr(), nbits, batch, _lib.stream_ptr())
    return state


def pick_device(tensors: Sequence[torch.Tensor]) -> torch.device:
    for t in tensors:
        if t.is_cuda:
            return t.device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    raise _lib.EngineError(
        "no CUDA device: the B200 engine has no CPU fallback (the numpy oracle under oracle/ is test "
        "infrastructure and is never used by the product path)"
    )


def build_gatebuf(tensors: Sequence[torch.Tensor], device: torch.device) -> torch.Tensor:
    flat = []
    for t in tensors:
        if t.dtype != torch.complex64:
            t = t.to(torch.complex64)
        if t.device != device:
            t = t.to(device)
        flat.append(t.reshape(-1))
    return torch.cat(flat) if flat else torch.zeros(1, dtype=torch.complex64, device=device)


_perm_cache: Dict[Any, torch.Tensor] = {}


def autograd_is_batched(t: torch.Tensor) -> bool:
    from . import autograd  # local import (autograd imports this module)

    return autograd.is_batched(t)


def assemble_gatebuf(gate_nodes: Sequence[Any], device: torch.device) -> torch.Tensor:
    """The gate buffer (every gate matrix, flat, in program order).  Deferred parametrised gates
    (`gates.LazyGate`) are built per family in one batched expression — a handful of torch ops and
    autograd nodes for the whole circuit instead of ~10 per gate — and one cached gather puts the
    pieces in program order."""
    pend = [hasattr(g, "pending") and g.pending() for g in gate_nodes]
    if not any(pend):
        return build_gatebuf([g.tensor for g in gate_nodes], device)
    fams: Dict[int, Tuple[Any, List[int]]] = {}
    eager: List[int] = []
    for i, g in enumerate(gate_nodes):
        if pend[i]:
            fams.setdefault(id(g._lazy.family), (g._lazy.family, []))[1].append(i)
        else:
            eager.append(i)
    pieces = [build_gatebuf([gate_nodes[i].tensor for i in eager], device)] if eager else []
    order: List[Tuple[int, int]] = [(i, int(gate_nodes[i].tensor.numel())) for i in eager]  # (gate, numel) in cat order
    for fam, idx in fams.values():
        with torch._C.DisableTorchFunction():  # (attribute reads only: no active torch-function mode has a say)
            ths = [gate_nodes[i]._lazy.current_theta() for i in idx]
            host_side = all(not t.is_cuda and not t.requires_grad and t.grad_fn is None and not autograd_is_batched(t)
                            for t in ths)  # fmt: skip
        if host_side:
            # host snapshots (gates._LazySpec): one upload for the whole family
            thetas = torch.tensor([float(t) for t in ths], dtype=torch.float32).to(device)
        else:
            thetas = torch.stack([t.reshape(()).to(device=device, dtype=torch.float32) for t in ths])
        pieces.append(fam.batched(thetas).reshape(-1))
        order.extend((i, fam.numel) for i in idx)
    cat = torch.cat(pieces) if len(pieces) > 1 else pieces[0]
    key = (tuple(order), str(device))
    perm = _perm_cache.get(key)
    if perm is None:
        start = {}
        off = 0
        for i, numel in order:
            start[i] = (off, numel)
            off += numel
        host = np.concatenate([np.arange(start[i][0], start[i][0] + start[i][1]) for i in range(len(gate_nodes))])
        perm = torch.from_numpy(host.astype(np.int64)).to(device)
        if len(_perm_cache) > 64:
            _perm_cache.clear()
        _perm_cache[key] = perm
    return cat.index_select(0, perm)


def run_circuit_network(nodes: Sequence[Any], output_edge_order: Sequence[Any]) -> torch.Tensor:
    """Forward statevector for a circuit-shaped network; returns a [2]*n tensor."""
    from . import autograd  # local import (autograd imports this module)

    n, init_node, gates = extract_gate_stream(nodes, output_edge_order)
    probe = [g[0]._lazy.theta if hasattr(g[0], "pending") and g[0].pending() else g[0].tensor for g in gates]
    with torch._C.DisableTorchFunction():  # (attribute reads on ~600 tensors)
        device = pick_device(probe + ([init_node.tensor] if init_node is not None else []))
        batched = any(autograd.is_batched(t) for t in probe)
        wants = [autograd.wants_grad(t) for t in probe] if torch.is_grad_enabled() else []
    structure = [(g[1], k, int(math.prod(g[0].shape))) for g, k in zip(gates, gate_kinds(gates))]
    cc = compile_circuit(n, structure, device, absorb_prefix=init_node is None and not batched)
    if torch.is_grad_enabled():
        for g in gates:
            if len(g[1]) > 2 and not (hasattr(g[0], "pending") and g[0].pending()) and g[0].tensor.requires_grad:
                raise _lib.EngineError(
                    f"gradient with respect to a {len(g[1])}-qubit gate matrix is not supported (the adjoint walk "
                    "differentiates 1- and 2-qubit gates; wider gates must be constants)"
                )
    init = None
    if init_node is not None:
        init = init_node.tensor.to(torch.complex64).to(device).reshape(-1)
    if torch.is_grad_enabled() and (any(wants) or (init is not None and autograd.wants_grad(init))):
        _require_unitary_for_adjoint(gates)
    gatebuf = assemble_gatebuf([g[0] for g in gates], device)
    const_mask = None
    if torch.is_grad_enabled() and len(gates) == len(cc.ops):
        # which gates are constants (no gradient wanted): the backward walk un-applies runs of them in fused passes
        const_mask = tuple(not w for w in wants)
    state = autograd.evolve(cc, gatebuf, init, const_mask)
    return state.reshape([2] * n)
