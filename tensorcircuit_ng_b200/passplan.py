"""
Gate-stream -> pass-program compiler for the fused statevector kernel
(csrc/pass_core.cuh documents the word layout consumed on the device).

Input  : the gates of a circuit in program order (recovered from the node list the
         contractor receives, tensorcircuit/cons.py:28-53 `_stable_id_`), each with its
         qubits, a structural kind (gates.py `_b200_kind`) and the offset of its matrix
         in the per-circuit gate buffer.
Output : a list of steps; a `PassStep` is ONE HBM pass applying many gates.

Scheduling is dependency based, not order based: a gate is *ready* once every earlier
gate on each of its qubits has been scheduled.  For a pass we pick T tile bits (the L
lowest index bits are always in, for coalescing) so that as many ready gates as
possible become applicable:

    diagonal 1q/2q gates ........ always applicable (their non-tile bits are CTA constants)
    dense 1q .................... its qubit must be a tile bit
    controlled-1q (1-2 ctrls) ... the *target* must be a tile bit, controls may be anywhere
    dense 2..4q ................. all qubits in the tile (shared-memory sub-pass)
    anything else ............... its own unfused launch (`GlobalStep`)

Inside a pass the scheduled gates are cut into register sub-passes of R = 5 tile bits
by the same greedy rule.  Pure host code (numpy only) — unit-tested on CPU against the
oracle through the kernel-logic emulator in tests/emu/.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

# ---- word layout: keep in sync with csrc/pass_core.cuh -------------------------------
PASS_MAGIC = 0x7CB20001
PASS_MAX_T = 13
PASS_R = 5
PASS_MAX_WORDS = 3072
H_MAGIC, H_T, H_L, H_NSUB, H_WORDS, H_NNONTILE, H_R = 0, 1, 2, 3, 4, 5, 6
H_TILEPOS, H_NONTILEPOS, HDR_WORDS = 8, 24, 80
S_NOPS, S_KIND, S_REGBITS, S_GRPBITS, S_WORDS, SUB_HDR_WORDS = 0, 1, 2, 10, 22, 24
SUB_REG, SUB_SMEM_DENSE = 0, 1
OP_WORDS = 8
OP_1Q, OP_C1Q, OP_DIAG1, OP_DIAG2, OP_DENSE = 1, 2, 3, 4, 5
QREF_BIT = 32
MIN_PASS_BITS = PASS_R + 5  # smallest state the tile kernel accepts


@dataclass
class GateOp:
    qubits: Tuple[int, ...]  # circuit qubit indices (0 = most significant)
    kind: Tuple[Any, ...]  # ("dense",) | ("diag",) | ("diagvec",) | ("ctrl", nctrl, pol)
    mat_off: int  # offset (complex elements) of the gate tensor in the gate buffer
    gid: int = -1  # position in the circuit (for gradients / debugging)

    @property
    def k(self) -> int:
        return len(self.qubits)

    @property
    def is_diag(self) -> bool:
        return self.kind[0] in ("diag", "diagvec")


@dataclass
class PassStep:
    program: np.ndarray  # int32 words
    tile_bits: int
    low_bits: int
    gate_ids: List[int]
    n_subpasses: int


@dataclass
class GlobalStep:
    gate: GateOp
    bitpos: Tuple[int, ...]


@dataclass
class Plan:
    nbits: int
    steps: List[Any]
    n_gates: int

    @property
    def n_passes(self) -> int:
        return sum(1 for s in self.steps if isinstance(s, PassStep))

    @property
    def n_launches(self) -> int:
        return len(self.steps)


def _applicable(g: GateOp, pos_of: Sequence[int], inside: Set[int], max_dense: int) -> Optional[str]:
    """How gate g can run when the bit positions `inside` are resident; None if it cannot."""
    kind = g.kind[0]
    if kind in ("diag", "diagvec"):
        return "diag" if g.k <= 2 else None
    if kind == "ctrl":
        nctrl = g.kind[1]
        if nctrl <= 2 and g.k == nctrl + 1:
            return "c1q" if pos_of[g.qubits[-1]] in inside else None
        kind = "dense"
    if g.k == 1:
        return "1q" if pos_of[g.qubits[0]] in inside else None
    if g.k <= max_dense and all(pos_of[q] in inside for q in g.qubits):
        return "dense"
    return None


class _Frontier:
    """Per-qubit program-order queues with O(1) ready checks."""

    def __init__(self, gates: Sequence[GateOp], nq: int) -> None:
        self.gates = gates
        self.queues: List[List[int]] = [[] for _ in range(nq)]
        for gi, g in enumerate(gates):
            for q in g.qubits:
                self.queues[q].append(gi)
        self.ptr = [0] * nq
        self.remaining = len(gates)

    def clone_ptr(self) -> List[int]:
        return list(self.ptr)

    def ready(self, gi: int, ptr: List[int]) -> bool:
        for q in self.gates[gi].qubits:
            qq = self.queues[q]
            if ptr[q] >= len(qq) or qq[ptr[q]] != gi:
                return False
        return True

    def simulate(self, inside: Set[int], pos_of: Sequence[int], max_dense: int, limit: int,
                 commit: bool = False, allow: Optional[Set[int]] = None) -> List[Tuple[int, str]]:  # fmt: skip
        """Greedily run every gate that is ready and applicable; returns [(gate index, how)]."""
        ptr = self.ptr if commit else self.clone_ptr()
        out: List[Tuple[int, str]] = []
        nq = len(self.queues)
        work = list(range(nq))
        inwork = [True] * nq
        while work and len(out) < limit:
            q = work.pop()
            inwork[q] = False
            while ptr[q] < len(self.queues[q]) and len(out) < limit:
                gi = self.queues[q][ptr[q]]
                if allow is not None and gi not in allow:
                    break
                if not self.ready(gi, ptr):
                    break
                how = _applicable(self.gates[gi], pos_of, inside, max_dense)
                if how is None:
                    break
                out.append((gi, how))
                for qq in self.gates[gi].qubits:
                    ptr[qq] += 1
                    if qq != q and not inwork[qq]:
                        work.append(qq)
                        inwork[qq] = True
        if commit:
            self.remaining -= len(out)
        return out


def _score(sched: List[Tuple[int, str]]) -> float:
    # diagonal gates are (nearly) free riders; what a pass buys is dense work
    return sum(1.0 if how != "diag" else 0.05 for _, how in sched)


def _grow(front: _Frontier, pos_of: Sequence[int], start: Set[int], size: int, candidates: Sequence[int],
          max_dense: int, limit: int, allow: Optional[Set[int]] = None) -> Set[int]:  # fmt: skip
    """Greedy growth of the resident bit set up to `size` bits."""
    inside = set(start)
    cands = [c for c in candidates if c not in inside]
    while len(inside) < size and cands:
        base = _score(front.simulate(inside, pos_of, max_dense, limit, allow=allow))
        best, best_gain = None, 0.0
        for c in cands:
            gain = _score(front.simulate(inside | {c}, pos_of, max_dense, limit, allow=allow)) - base
            if gain > best_gain + 1e-9:
                best, best_gain = c, gain
        if best is None:
            # nothing helps on its own: look for a pair (dense 2q gates need both qubits resident)
            found = False
            for gi_q in range(len(front.queues)):
                qq = front.queues[gi_q]
                if front.ptr[gi_q] >= len(qq):
                    continue
                gi = qq[front.ptr[gi_q]]
                if allow is not None and gi not in allow:
                    continue
                g = front.gates[gi]
                need = [pos_of[q] for q in g.qubits if pos_of[q] not in inside]
                if g.is_diag or not need or len(inside) + len(need) > size:
                    continue
                if not all(p in cands for p in need) or not front.ready(gi, front.ptr):
                    continue
                if g.kind[0] == "dense" and g.k > max_dense:
                    continue
                for p in need:
                    inside.add(p)
                    cands.remove(p)
                found = True
                break
            if not found:
                break
        else:
            inside.add(best)
            cands.remove(best)
    # fill up with arbitrary candidates (keeps the kernel's fixed geometry)
    for c in cands:
        if len(inside) >= size:
            break
        inside.add(c)
    return inside


def _order_group_bits(nonreg: List[int]) -> List[int]:
    """Order the non-register tile bits so the 4 lowest thread-id bits land on distinct
    residues mod 4 (conflict-free under the XOR-fold swizzle of pass_core.cuh::swz)."""
    rest = sorted(nonreg)
    first: List[int] = []
    used = set()
    for b in rest:
        if b % 4 not in used and len(first) < 4:
            first.append(b)
            used.add(b % 4)
    tail = [b for b in rest if b not in first]
    return first + tail


def _encode_pass(nbits: int, tile_pos: List[int], L: int, subpasses: List[Dict[str, Any]]) -> np.ndarray:
    T = len(tile_pos)
    words: List[int] = [0] * HDR_WORDS
    words[H_MAGIC] = PASS_MAGIC
    words[H_T] = T
    words[H_L] = L
    words[H_NSUB] = len(subpasses)
    words[H_R] = PASS_R
    nontile = [p for p in range(nbits) if p not in set(tile_pos)]
    words[H_NNONTILE] = len(nontile)
    assert len(nontile) <= 56 and T <= 16
    for i, p in enumerate(tile_pos):
        words[H_TILEPOS + i] = p
    for i, p in enumerate(nontile):
        words[H_NONTILEPOS + i] = p
    for sp in subpasses:
        hdr = [0] * SUB_HDR_WORDS
        hdr[S_NOPS] = len(sp["ops"])
        hdr[S_KIND] = sp["kind"]
        for j, b in enumerate(sp.get("reg", [])):
            hdr[S_REGBITS + j] = b
        for j, b in enumerate(sp.get("grp", [])):
            hdr[S_GRPBITS + j] = b
        hdr[S_WORDS] = SUB_HDR_WORDS + OP_WORDS * len(sp["ops"])
        words.extend(hdr)
        for op in sp["ops"]:
            assert len(op) == OP_WORDS
            words.extend(op)
    words[H_WORDS] = len(words)
    return np.asarray(words, dtype=np.int32)


def _dense_dim(g: GateOp) -> int:
    return 1 << g.k


def compile_plan(gates: Sequence[GateOp], nqubits: int, *, nbits_local: Optional[int] = None,
                 tile_bits: int = PASS_MAX_T, low_bits: int = 4, max_ops_per_pass: int = 300,
                 lookahead: int = 512) -> Plan:  # fmt: skip
    """Compile a gate stream.  `nbits_local` < nqubits describes a sharded state whose top
    (nqubits - nbits_local) qubits are global: only diagonal gates / controls may touch them."""
    nbits = nqubits if nbits_local is None else nbits_local
    pos_of = [nqubits - 1 - q for q in range(nqubits)]  # qubit -> flat bit position
    for gi, g in enumerate(gates):
        if g.gid < 0:
            g.gid = gi
    steps: List[Any] = []
    front = _Frontier(gates, nqubits)
    use_pass = nbits >= MIN_PASS_BITS
    T = min(tile_bits, nbits, PASS_MAX_T)
    L = min(low_bits, T)
    local_positions = list(range(nbits))

    def emit_global(gi: int) -> None:
        g = gates[gi]
        bp = tuple(pos_of[q] for q in g.qubits)
        if not g.is_diag and any(p >= nbits for p in bp):
            raise ValueError(
                f"gate {g.gid} on qubits {g.qubits} needs a dense operation on a global qubit; "
                "swap it into the local register first (sharded.py)"
            )
        steps.append(GlobalStep(g, bp))
        for q in g.qubits:
            front.ptr[q] += 1
        front.remaining -= 1

    while front.remaining > 0:
        if not use_pass:
            # tiny registers: literal program order, one launch per gate
            nxt = min(front.queues[q][front.ptr[q]] for q in range(nqubits) if front.ptr[q] < len(front.queues[q]))
            # program order is a valid topological order
            emit_global(nxt)
            continue
        start = set(range(L))
        inside = _grow(front, pos_of, start, T, local_positions, 4, lookahead)
        sched = front.simulate(inside, pos_of, 4, max_ops_per_pass)
        if _score(sched) < 0.5 and all(how == "diag" for _, how in sched):
            # no dense work can be unlocked by any tile: the earliest blocked gate goes unfused
            blocked = [
                front.queues[q][front.ptr[q]]
                for q in range(nqubits)
                if front.ptr[q] < len(front.queues[q])
            ]
            blocked = [gi for gi in blocked if front.ready(gi, front.ptr)]
            hard = [gi for gi in blocked if _applicable(gates[gi], pos_of, set(local_positions), 4) is None]
            if hard and not sched:
                emit_global(min(hard))
                continue
            if not sched:
                raise RuntimeError("pass planner stalled")  # pragma: no cover
        sched = front.simulate(inside, pos_of, 4, max_ops_per_pass, commit=True)
        steps.append(_build_pass(gates, sched, nbits, pos_of, sorted(inside), L))
    return Plan(nbits=nbits, steps=steps, n_gates=len(gates))


def _build_pass(gates: Sequence[GateOp], sched: List[Tuple[int, str]], nbits: int, pos_of: Sequence[int],
                tile_pos: List[int], L: int) -> PassStep:  # fmt: skip
    T = len(tile_pos)
    tbit_of_pos = {p: i for i, p in enumerate(tile_pos)}
    nq = len(pos_of)
    sub_gates = [gates[gi] for gi, _ in sched]
    allow = {i for i in range(len(sub_gates))}
    # local frontier over the scheduled gates only (indices into sub_gates)
    local = _Frontier(sub_gates, nq)
    subpasses: List[Dict[str, Any]] = []
    tile_set = set(tile_pos)
    while local.remaining > 0:
        # is the next ready gate a dense k>=2 gate?  then it is its own shared-memory sub-pass
        inside = _grow(local, pos_of, set(), PASS_R, tile_pos, 1, 1 << 30)
        run = local.simulate(inside, pos_of, 1, 1 << 30)
        if not run:
            # only dense multi-qubit gates are ready
            cand = [
                local.queues[q][local.ptr[q]] for q in range(nq) if local.ptr[q] < len(local.queues[q])
            ]
            cand = sorted({gi for gi in cand if local.ready(gi, local.ptr)})
            gi = cand[0]
            g = sub_gates[gi]
            assert g.k >= 2 and all(pos_of[q] in tile_set for q in g.qubits), "planner invariant"
            tb = [tbit_of_pos[pos_of[q]] for q in g.qubits] + [0] * 5
            op = [OP_DENSE, tb[0], tb[1], g.mat_off, g.k, tb[2], tb[3], tb[4]]
            subpasses.append({"kind": SUB_SMEM_DENSE, "ops": [op]})
            for q in g.qubits:
                local.ptr[q] += 1
            local.remaining -= 1
            continue
        run = local.simulate(inside, pos_of, 1, 1 << 30, commit=True)
        reg_pos = sorted(inside)  # flat bit positions held in registers
        reg_tb = [tbit_of_pos[p] for p in reg_pos]
        reg_idx = {p: j for j, p in enumerate(reg_pos)}
        grp = _order_group_bits([t for t in range(T) if t not in set(reg_tb)])

        def qref(q: int) -> int:
            p = pos_of[q]
            return reg_idx[p] if p in reg_idx else QREF_BIT + p

        ops: List[List[int]] = []
        for gi, how in run:
            g = sub_gates[gi]
            D = _dense_dim(g)
            if how == "1q":
                ops.append([OP_1Q, reg_idx[pos_of[g.qubits[0]]], 0, g.mat_off, 0, 2, 0, 0])
            elif how == "c1q":
                nctrl, pol = g.kind[1], g.kind[2]
                wants = [(pol >> i) & 1 for i in range(nctrl)]
                polval = 0
                for w in wants:
                    polval = (polval << 1) | w
                mat = g.mat_off + (polval * 2) * D + polval * 2
                qc0 = qref(g.qubits[0])
                qc1 = qref(g.qubits[1]) if nctrl == 2 else -1
                polbits = wants[0] | ((wants[1] << 1) if nctrl == 2 else 0)
                ops.append([OP_C1Q, qc0, reg_idx[pos_of[g.qubits[-1]]], mat, polbits, D, qc1, 0])
            elif how == "diag":
                packed = g.kind[0] == "diagvec"
                if g.k == 1:
                    ops.append([OP_DIAG1, qref(g.qubits[0]), 0, g.mat_off, 0, 1 if packed else 3, 0, 0])
                else:
                    ops.append(
                        [OP_DIAG2, qref(g.qubits[0]), qref(g.qubits[1]), g.mat_off, 0, 1 if packed else 5, 0, 0]
                    )
            else:  # pragma: no cover
                raise AssertionError(how)
        subpasses.append({"kind": SUB_REG, "reg": reg_tb, "grp": grp, "ops": ops})
    program = _encode_pass(nbits, tile_pos, L, subpasses)
    if len(program) > PASS_MAX_WORDS:
        raise RuntimeError(f"pass program too large ({len(program)} words)")
    return PassStep(program=program, tile_bits=T, low_bits=L, gate_ids=[gates[gi].gid for gi, _ in sched],
                    n_subpasses=len(subpasses))  # fmt: skip
