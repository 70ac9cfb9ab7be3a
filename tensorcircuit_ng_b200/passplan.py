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

    diagonal 1q/2q gates ........ applicable when ONE of their qubits is a tile bit (the partner may
                                  be a CTA constant).  They are scheduled lazily: a diagonal gate is
                                  attached to a round in which one of its qubits is a register bit,
                                  and otherwise left for a later pass
    dense 1q .................... its qubit must be a tile bit
    controlled-1q (1-2 ctrls) ... the *target* must be a tile bit, controls may be anywhere
    dense 2..4q ................. all qubits in the tile (shared-memory sub-pass)
    anything else ............... its own unfused launch (`GlobalStep`)

Inside a pass the scheduled gates are cut into register sub-passes of R = 5 tile bits
(slot 0 is always tile bit 0, so shared memory is accessed in 16-byte amplitude pairs) by the
same greedy rule.  Pure host code (numpy only) — unit-tested on CPU against the
oracle through the kernel-logic emulator in tests/emu/.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

# ---- word layout: keep in sync with csrc/pass_core.cuh -------------------------------
PASS_MAGIC = 0x7CB20004
PASS_MAX_T = 13
PASS_R = 5
PASS_MAX_WORDS = 6144
PASS_MAX_SUB = 16
H_MAGIC, H_T, H_L, H_NSUB, H_WORDS, H_NNONTILE, H_R, H_NFILL = 0, 1, 2, 3, 4, 5, 6, 7
H_NPOOL, H_POOLSIZE, H_NFILL_STATIC, H_NFILL_SHORT_END = 64, 65, 66, 67
FILL_SHORT = 4
PASS_MAX_POOL = 1536
H_TILEPOS, H_NONTILEPOS, HDR_WORDS = 8, 24, 80
S_NROUNDS, S_KIND, S_REGBITS, S_GRPBITS, S_WORDS, SUB_HDR_WORDS = 0, 1, 2, 10, 22, 24
SUB_REG, SUB_SMEM_DENSE = 0, 1
RD_FLAGS, RD_NRR, RD_WORDS, RD_NRJ, RD_CTRL, RD_M, RD_F, RD_FIXED = 0, 1, 2, 3, 8, 20, 60, 80
RD_GENERAL = 1 << 30
TT_A, TT_B, TT_W, TT_WORDS = 0, 1, 4, 12
OP_WORDS = 16
OP_DENSE = 5
FK_PAIR, FK_TABLE, FK_MATRIX = 1, 2, 3
FF_D1, FF_D2_FIRST, FF_D2_SECOND, FF_T, FF_T_SWAP, FF_M, FF_MD = 0, 1, 2, 5, 6, 7, 8
MIN_PASS_BITS = PASS_R + 5  # smallest state the tile kernel accepts
_ONE = int(np.float32(1.0).view(np.int32))


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
    pool_elems: int = 0  # complex elements of the shared-memory gate pool (sizes the launch)


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
        # lazily scheduled: one of its qubits must be resident (the gate becomes per-thread data of
        # the round that owns that bit, pass_core.cuh); a partner outside is a CTA / thread constant
        if g.k <= 2:
            nin = sum(1 for q in g.qubits if pos_of[q] in inside)
            if nin:
                return "diag2" if nin == 2 else "diag"
        return None
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

    def heads(self) -> List[int]:
        """Ready gates (head of every queue they sit in), ascending."""
        out = set()
        for q, qq in enumerate(self.queues):
            if self.ptr[q] < len(qq) and self.ready(qq[self.ptr[q]], self.ptr):
                out.add(qq[self.ptr[q]])
        return sorted(out)

    def commit(self, done: Sequence[int]) -> None:
        """Mark `done` (a per-qubit prefix-closed set of pending gates) as executed."""
        dset = set(done)
        for gi in done:
            for q in self.gates[gi].qubits:
                assert self.queues[q][self.ptr[q]] in dset, "planner invariant: executed set is not prefix closed"
        cnt = [0] * len(self.queues)
        for gi in done:
            for q in self.gates[gi].qubits:
                cnt[q] += 1
        for q, c in enumerate(cnt):
            self.ptr[q] += c
        self.remaining -= len(done)

    def simulate(self, inside: Set[int], pos_of: Sequence[int], max_dense: int, limit: int,
                 commit: bool = False, allow: Optional[Set[int]] = None,
                 pred: Optional[Any] = None) -> List[Tuple[int, str]]:  # fmt: skip
        """Greedily run every gate that is ready and applicable; returns [(gate index, how)].
        `pred(gate) -> how | None` replaces the tile applicability rule (sharded.py)."""
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
                how = _applicable(self.gates[gi], pos_of, inside, max_dense) if pred is None else pred(self.gates[gi])
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


_RR_PENALTY = -0.75


def _score(sched: List[Tuple[int, str]], rr: float = 0.05) -> float:
    """What a resident bit set buys is dense work; diagonal gates are (nearly) free riders — except,
    at the register level (rr < 0), a 2q diagonal with BOTH qubits in registers, which costs a
    per-amplitude multiply instead of a per-thread one."""
    tot = 0.0
    for _, how in sched:
        tot += 0.05 if how == "diag" else (rr if how == "diag2" else 1.0)
    return tot


def _grow(front: _Frontier, pos_of: Sequence[int], start: Set[int], size: int, candidates: Sequence[int],
          max_dense: int, limit: int, allow: Optional[Set[int]] = None, rr: float = 0.05,
          fill: bool = True) -> Set[int]:  # fmt: skip
    """Greedy growth of the resident bit set up to `size` bits."""
    inside = set(start)
    cands = [c for c in candidates if c not in inside]
    while len(inside) < size and cands:
        base = _score(front.simulate(inside, pos_of, max_dense, limit, allow=allow), rr)
        best, best_gain = None, 0.0
        for c in cands:
            gain = _score(front.simulate(inside | {c}, pos_of, max_dense, limit, allow=allow), rr) - base
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
    # fill up with candidates that unlock nothing (keeps the kernel's fixed geometry)
    if fill:
        base = front.simulate(inside, pos_of, max_dense, limit, allow=allow)
        for c in cands:
            if len(inside) >= size:
                break
            if len(front.simulate(inside | {c}, pos_of, max_dense, limit, allow=allow)) == len(base):
                inside.add(c)
        for c in cands:
            if len(inside) >= size:
                break
            inside.add(c)
    return inside


def _order_group_bits(nonreg: List[int]) -> List[int]:
    """Order the non-register tile bits so the 3 lowest thread-id bits (a quarter warp = one
    LDS.128 wavefront) land on the 3 distinct chunk classes of pass_core.cuh::swz, (b - 1) mod 3."""
    rest = sorted(nonreg)
    first: List[int] = []
    used = set()
    for b in rest:
        if (b - 1) % 3 not in used and len(first) < 3:
            first.append(b)
            used.add((b - 1) % 3)
    tail = [b for b in rest if b not in first]
    return first + tail


def terminal_diagonals(gates: Sequence[GateOp], nq: int) -> Set[int]:
    """Diagonal gates with no later non-diagonal gate on any of their qubits (never worth deferring)."""
    later_dense = [False] * nq
    out: Set[int] = set()
    for gi in range(len(gates) - 1, -1, -1):
        g = gates[gi]
        if g.is_diag:
            if not any(later_dense[q] for q in g.qubits):
                out.add(gi)
        else:
            for q in g.qubits:
                later_dense[q] = True
    return out


def compile_plan(gates: Sequence[GateOp], nqubits: int, *, nbits_local: Optional[int] = None,
                 tile_bits: int = 12, low_bits: int = 3, max_ops_per_pass: int = 200,
                 lookahead: int = 512, pos_of: Optional[Sequence[int]] = None) -> Plan:  # fmt: skip
    """Compile a gate stream.  `nbits_local` < nqubits describes a sharded state: flat-index positions
    >= nbits_local are global (rank bits): only diagonal gates / controls may touch them.
    `pos_of[q]` = flat-index bit position of qubit q (default: qubit 0 is the most significant bit,
    tensorcircuit/circuit.py:711-719; the sharded scheduler passes its current qubit layout)."""
    nbits = nqubits if nbits_local is None else nbits_local
    pos_of = [nqubits - 1 - q for q in range(nqubits)] if pos_of is None else list(pos_of)
    for gi, g in enumerate(gates):
        if g.gid < 0:
            g.gid = gi
    steps: List[Any] = []
    front = _Frontier(gates, nqubits)
    terminal = terminal_diagonals(gates, nqubits)
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
        front.commit([gi])

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
        if not sched:
            # nothing can run inside any tile: the earliest ready gate goes unfused (k >= 5 dense
            # gates, diagonal gates on global qubits only, k >= 3 diagonals)
            heads = front.heads()
            if not heads:
                raise RuntimeError("pass planner stalled")  # pragma: no cover
            emit_global(heads[0])
            continue
        # the program must fit the kernel's shared-memory budget: shrink the pass until it does
        limit = max_ops_per_pass
        step = None
        while True:
            sched = front.simulate(inside, pos_of, 4, limit)
            try:
                step, done = _build_pass(gates, sched, nbits, pos_of, sorted(inside), L, terminal)
                if not done:  # everything was deferred: apply the diagonal gates right here
                    step, done = _build_pass(gates, sched, nbits, pos_of, sorted(inside), L, None)
                break
            except _ProgramTooLarge:
                if len(sched) <= 1:
                    raise
                limit = max(1, int(len(sched) * 0.7))
        if not done:
            heads = front.heads()
            emit_global(heads[0])
            continue
        front.commit(done)
        steps.append(step)
    return Plan(nbits=nbits, steps=steps, n_gates=len(gates))


class _ProgramTooLarge(RuntimeError):
    pass


class _Pool:
    """Gate tensors a pass needs, packed for the CTA's shared-memory pool (pass_kernel.cu)."""

    def __init__(self) -> None:
        self.entries: Dict[int, Tuple[int, int]] = {}  # gate mat_off -> (numel, pool offset)
        self.size = 0

    def ref(self, g: GateOp, delta: int = 0) -> int:
        if g.mat_off not in self.entries:
            numel = (1 << g.k) if g.kind[0] == "diagvec" else (1 << (2 * g.k))
            self.entries[g.mat_off] = (numel, self.size)
            self.size += numel
        return self.entries[g.mat_off][1] + delta

    def table(self) -> List[int]:
        out: List[int] = []
        for mat_off, (numel, poff) in self.entries.items():
            out += [mat_off, numel, poff]
        return out


class _Round:
    def __init__(self) -> None:
        self.gate: List[Optional[Dict[str, Any]]] = [None] * PASS_R
        self.ctrl_regs: Set[int] = set()
        # diagonal part (applied BEFORE the gates of the round)
        self.pair_src: List[List[int]] = [[] for _ in range(PASS_R)]  # fill sources of slot J's factor
        self.rj: List[List[Tuple[int, GateOp, int]]] = [[] for _ in range(PASS_R)]  # (partner tile bit, gate, form)
        self.rr: List[Tuple[int, int, GateOp]] = []  # (slot of first qubit, slot of second qubit, gate)

    def has_factor(self, j: int) -> bool:
        return bool(self.pair_src[j]) or bool(self.rj[j])

    def empty(self) -> bool:
        return (all(x is None for x in self.gate) and not self.rr
                and not any(self.has_factor(j) for j in range(PASS_R)))  # fmt: skip


def _diag_stride(g: GateOp) -> int:
    return 1 if g.kind[0] == "diagvec" else (1 << g.k) + 1


def _build_rounds(sub_gates: Sequence[GateOp], run: List[Tuple[int, str]], pos_of: Sequence[int],
                  reg_idx: Dict[int, int], tbit_of_pos: Dict[int, int], pool: _Pool) -> List[_Round]:  # fmt: skip
    """List-schedule the ordered gate list of one register sub-pass into rounds (pass_core.cuh: a
    round is, per register slot, an optional diagonal factor followed by at most one fused 2x2).

    Diagonal gates commute with everything except a non-diagonal gate on one of their own qubits,
    so they stay *pending* on their register slot(s) until the next gate on such a slot is placed and
    then ride in that gate's round as its factor (no extra rounds, no factor-only multiplies); what
    is still pending at the end goes into one trailing round."""
    rounds: List[_Round] = []
    fused_block = [False] * PASS_R  # something that does not commute came after the slot's latest gate
    last_gate = [-1] * PASS_R  # round of the latest gate on slot j
    floor = [0] * PASS_R  # earliest round the next gate on slot j may take
    pend: List[List[Tuple[str, Any]]] = [[] for _ in range(PASS_R)]
    pend_rr: List[Tuple[int, int, GateOp]] = []

    def ensure(r: int) -> _Round:
        while len(rounds) <= r:
            rounds.append(_Round())
        return rounds[r]

    def flush(j: int, r: int) -> None:
        """Move everything pending on slot j into the diagonal part of round r."""
        rnd = ensure(r)
        for kind, data in pend[j]:
            if kind == "pair":
                rnd.pair_src[j].extend(data)
            else:
                rnd.rj[j].append(data)
        pend[j].clear()
        keep = []
        for ja, jb, g in pend_rr:
            if j in (ja, jb):
                rnd.rr.append((ja, jb, g))
                other = jb if j == ja else ja
                floor[other] = max(floor[other], r)  # diagonal part of round r precedes its gates
                fused_block[other] = True  # a later gate on `other` must not join its earlier product
            else:
                keep.append((ja, jb, g))
        pend_rr[:] = keep

    def rr_floor(j: int) -> int:
        r = 0
        for ja, jb, _ in pend_rr:
            if j in (ja, jb):
                r = max(r, last_gate[jb if j == ja else ja] + 1)
        return r

    def can_fuse(j: int) -> bool:
        r = last_gate[j]
        return (r >= 0 and rounds[r].gate[j]["ctrl"] is None and not pend[j] and floor[j] <= r + 1
                and not any(j in (ja, jb) for ja, jb, _ in pend_rr) and j not in rounds[r].ctrl_regs
                and not fused_block[j])  # fmt: skip

    for gi, how in run:
        g = sub_gates[gi]
        if how in ("diag", "diag2"):
            st = _diag_stride(g)
            ps = [pos_of[q] for q in g.qubits]
            slots = [reg_idx.get(p) for p in ps]
            if g.k == 1:
                j = slots[0]
                assert j is not None
                if can_fuse(j):
                    # 1q diagonal right after dense gate(s) on the same slot: same 2x2 product
                    rounds[last_gate[j]].gate[j]["fusion"].append((g, 0, FF_MD, st))
                else:
                    pend[j].append(("pair", [pool.ref(g), FF_D1 | (st << 8), 0, 0]))
                continue
            if slots[0] is not None and slots[1] is not None:
                pend_rr.append((slots[0], slots[1], g))
                continue
            own = 0 if slots[0] is not None else 1  # which gate qubit is the register bit
            j = slots[own]
            assert j is not None
            pp = ps[1 - own]
            if pp in tbit_of_pos:  # thread-constant partner: one LDS.128 lookup per thread
                # W = [x_B=0: (d[x_A=0], d[x_A=1])], [x_B=1: ...]; A is the register bit
                form = FF_T_SWAP if own == 0 else FF_T
                pend[j].append(("rj", (tbit_of_pos[pp], g, form)))
            else:  # CTA-constant partner: resolved per tile by the prologue
                form = FF_D2_FIRST if own == 0 else FF_D2_SECOND
                pend[j].append(("pair", [pool.ref(g), form | (st << 8), pp, 0]))
        elif how == "1q":
            j = reg_idx[pos_of[g.qubits[0]]]
            if can_fuse(j):
                rounds[last_gate[j]].gate[j]["fusion"].append((g, 0, FF_M, 2))
                continue
            r = max(floor[j], rr_floor(j))
            rnd = ensure(r)
            assert rnd.gate[j] is None
            rnd.gate[j] = {"fusion": [(g, 0, FF_M, 2)], "ctrl": None}
            flush(j, r)
            last_gate[j], floor[j], fused_block[j] = r, r + 1, False
        elif how == "c1q":
            nctrl, pol = g.kind[1], g.kind[2]
            j = reg_idx[pos_of[g.qubits[-1]]]
            cmask = cwant = 0
            tcs: List[int] = []
            regs: Set[int] = set()
            for ci in range(nctrl):
                want = (pol >> ci) & 1
                p = pos_of[g.qubits[ci]]
                if p in reg_idx:
                    k = reg_idx[p]
                    cmask |= 1 << k
                    cwant |= want << k
                    regs.add(k)
                elif p in tbit_of_pos:
                    tcs.append(tbit_of_pos[p] | 0x80 | (want << 8))
                else:
                    tcs.append(p | (want << 8))
            r = max(floor[j], rr_floor(j))
            for k in regs:  # a round applies its gates in slot order, not program order
                r = max(r, last_gate[k] + 1)
            rnd = ensure(r)
            assert rnd.gate[j] is None
            D = 1 << g.k
            polval = 0
            for ci in range(nctrl):
                polval = (polval << 1) | ((pol >> ci) & 1)
            delta = (polval * 2) * D + polval * 2  # top-left of the active 2x2 block
            rnd.gate[j] = {"fusion": [(g, delta, FF_M, D)], "ctrl": {"cmask": cmask, "cwant": cwant, "tcs": tcs}}
            rnd.ctrl_regs |= regs
            flush(j, r)
            last_gate[j], floor[j], fused_block[j] = r, r + 1, False
            for k in regs:
                floor[k] = max(floor[k], r + 1)
                fused_block[k] = True
        else:  # pragma: no cover
            raise AssertionError(how)
    # whatever is still pending: one trailing diagonal part per slot, after the slot's last gate
    for j in range(PASS_R):
        if pend[j] or any(j in (ja, jb) for ja, jb, _ in pend_rr):
            flush(j, max(last_gate[j] + 1, floor[j], rr_floor(j)))
    return [r for r in rounds if not r.empty()]


def _emit_round(words: List[int], fills: List[List[int]], rnd: _Round, pool: _Pool) -> None:
    base = len(words)
    assert base % 4 == 0
    rec = [0] * RD_FIXED
    flags = 0
    for j in range(PASS_R):
        rec[RD_M + 8 * j] = _ONE  # identity defaults (overwritten by the device prologue)
        rec[RD_M + 8 * j + 6] = _ONE
        rec[RD_F + 4 * j] = _ONE
        rec[RD_F + 4 * j + 2] = _ONE
    for j, gt in enumerate(rnd.gate):
        if gt is None:
            continue
        flags |= 1 << j
        fills.append([base + RD_M + 8 * j, FK_MATRIX, len(gt["fusion"]), 0]
                     + [w for (fg, delta, form, st) in gt["fusion"]
                        for w in (pool.ref(fg, delta), form | (st << 8), 0, 0)])  # fmt: skip
        c = gt["ctrl"]
        if c is not None:
            if c["cmask"]:
                flags |= 1 << (8 + j)
            tcs = c["tcs"]
            w0 = c["cmask"] | (c["cwant"] << 8) | (len(tcs) << 16)
            w1 = 0
            if len(tcs) > 0:
                w1 |= tcs[0]
            if len(tcs) > 1:
                w1 |= tcs[1] << 16
            rec[RD_CTRL + 2 * j] = w0
            rec[RD_CTRL + 2 * j + 1] = w1
    tables: List[Tuple[int, int, GateOp, int]] = []
    for j in range(PASS_R):
        if rnd.has_factor(j):
            flags |= 1 << (16 + j)
        if rnd.pair_src[j]:
            fills.append([base + RD_F + 4 * j, FK_PAIR, len(rnd.pair_src[j]) // 4, 0] + rnd.pair_src[j])
        rec[RD_NRJ + j] = len(rnd.rj[j])
        for tb, g, form in rnd.rj[j]:
            tables.append((j, tb, g, form))
    rec[RD_NRR] = len(rnd.rr)
    for ja, jb, g in rnd.rr:  # W[2 * x_A + x_B] with A < B (pass_core.cuh dispatches on the slot pair)
        tables.append((ja, jb, g, FF_T) if ja < jb else (jb, ja, g, FF_T_SWAP))
    if any(gt is not None and gt["ctrl"] is not None for gt in rnd.gate) or any(
            rnd.has_factor(j) and rnd.gate[j] is None for j in range(PASS_R)):  # fmt: skip
        flags |= RD_GENERAL  # pass_core.cuh: everything else takes the compact fast path
    rec[RD_FLAGS] = flags
    rec[RD_WORDS] = RD_FIXED + TT_WORDS * len(tables)
    words.extend(rec)
    for a, b, g, form in tables:
        ent = [0] * TT_WORDS
        ent[TT_A], ent[TT_B] = a, b
        for i in range(4):
            ent[TT_W + 2 * i] = _ONE
        fills.append([len(words) + TT_W, FK_TABLE, 1, 0, pool.ref(g), form | (_diag_stride(g) << 8), 0, 0])
        words.extend(ent)


def _needed_prefix(sub_gates: Sequence[GateOp], run: List[Tuple[int, str]], keep_always: Optional[Set[int]],
                   gids: Sequence[int]) -> List[Tuple[int, str]]:  # fmt: skip
    """Drop the diagonal gates of `run` that nothing later in the run depends on (they stay pending
    and ride along, for free, with the next gate on one of their qubits — possibly in a later pass).
    The kept set is closed under "earlier gate on a shared qubit", so it is a valid schedule."""
    if keep_always is None:
        return run
    needed: Set[int] = set()
    keep = [False] * len(run)
    for i in range(len(run) - 1, -1, -1):
        gi, how = run[i]
        g = sub_gates[gi]
        if how != "diag" or gids[gi] in keep_always or any(q in needed for q in g.qubits):
            keep[i] = True
            needed.update(g.qubits)
    return [x for x, k in zip(run, keep) if k]


def _build_pass(gates: Sequence[GateOp], sched: List[Tuple[int, str]], nbits: int, pos_of: Sequence[int],
                tile_pos: List[int], L: int, terminal: Optional[Set[int]]) -> Tuple[PassStep, List[int]]:  # fmt: skip
    """Returns the pass and the indices (into `gates`) of the gates it executes.  `terminal` = the
    diagonal gates that must not be deferred (None: defer nothing)."""
    T = len(tile_pos)
    tbit_of_pos = {p: i for i, p in enumerate(tile_pos)}
    assert tile_pos[0] == 0, "tile bit 0 must be flat bit 0 (register slot 0)"
    nq = len(pos_of)
    sub_gates = [gates[gi] for gi, _ in sched]
    gids = [gi for gi, _ in sched]
    local = _Frontier(sub_gates, nq)  # frontier over the scheduled gates only
    tile_set = set(tile_pos)
    done_local: List[int] = []

    words: List[int] = [0] * HDR_WORDS
    words[H_MAGIC] = PASS_MAGIC
    words[H_T] = T
    words[H_L] = L
    words[H_R] = PASS_R
    nontile = [p for p in range(nbits) if p not in tile_set]
    words[H_NNONTILE] = len(nontile)
    assert len(nontile) <= 40 and T <= 16
    for i, p in enumerate(tile_pos):
        words[H_TILEPOS + i] = p
    for i, p in enumerate(nontile):
        words[H_NONTILEPOS + i] = p
    fills: List[List[int]] = []
    pool = _Pool()
    nsub = 0
    while local.remaining > 0 and nsub < PASS_MAX_SUB:
        inside = _grow(local, pos_of, {0}, PASS_R, tile_pos, 1, 1 << 30, rr=_RR_PENALTY)
        run = _needed_prefix(sub_gates, local.simulate(inside, pos_of, 1, 1 << 30), terminal, gids)
        if not run:
            # no register work is due: a dense multi-qubit gate, if one is ready, takes a
            # shared-memory sub-pass; diagonal gates that block one are applied right away
            heads = local.heads()
            dense = [gi for gi in heads if not sub_gates[gi].is_diag]
            if dense:
                g = sub_gates[dense[0]]
                assert g.k >= 2 and all(pos_of[q] in tile_set for q in g.qubits), "planner invariant"
                tb = [tbit_of_pos[pos_of[q]] for q in g.qubits] + [0] * 4
                hdr = [0] * SUB_HDR_WORDS
                hdr[S_NROUNDS], hdr[S_KIND], hdr[S_WORDS] = 1, SUB_SMEM_DENSE, SUB_HDR_WORDS + OP_WORDS
                words.extend(hdr)
                words.extend([OP_DENSE, tb[0], tb[1], g.mat_off, g.k, tb[2], tb[3], 0] + [0] * 8)
                nsub += 1
                local.commit([dense[0]])
                done_local.append(dense[0])
                continue
            blocking = [gi for gi in heads if any(
                any(not sub_gates[x].is_diag for x in local.queues[q][local.ptr[q] + 1:])
                for q in sub_gates[gi].qubits)]  # fmt: skip
            if not blocking:
                break  # only deferred diagonal gates are left: a later pass takes them
            inside = _grow(local, pos_of, {0}, PASS_R, tile_pos, 1, 1 << 30, allow=set(blocking))
            run = local.simulate(inside, pos_of, 1, 1 << 30, allow=set(blocking))
            if not run:
                break
        local.commit([gi for gi, _ in run])
        done_local.extend(gi for gi, _ in run)
        reg_pos = [0] + sorted(p for p in inside if p != 0)  # slot 0 = flat bit 0 = tile bit 0
        reg_tb = [tbit_of_pos[p] for p in reg_pos]
        reg_idx = {p: j for j, p in enumerate(reg_pos)}
        grp = _order_group_bits([t for t in range(T) if t not in set(reg_tb)])
        rounds = _build_rounds(sub_gates, run, pos_of, reg_idx, tbit_of_pos, pool)
        hdr_at = len(words)
        hdr = [0] * SUB_HDR_WORDS
        hdr[S_NROUNDS], hdr[S_KIND] = len(rounds), SUB_REG
        for j, b in enumerate(reg_tb):
            hdr[S_REGBITS + j] = b
        for j, b in enumerate(grp):
            hdr[S_GRPBITS + j] = b
        words.extend(hdr)
        for rnd in rounds:
            _emit_round(words, fills, rnd, pool)
        words[hdr_at + S_WORDS] = len(words) - hdr_at
        nsub += 1
    words[H_NSUB] = nsub
    # fill records [static | dynamic short | dynamic long], then the table of their offsets
    def _is_static(f: List[int]) -> bool:  # no source looks at CTA-constant bits
        return all((f[4 + 4 * i + 1] & 0xFF) not in (FF_D2_FIRST, FF_D2_SECOND) for i in range(f[2]))

    stat = [f for f in fills if _is_static(f)]
    dyn_short = [f for f in fills if not _is_static(f) and f[2] <= FILL_SHORT]
    dyn_long = [f for f in fills if not _is_static(f) and f[2] > FILL_SHORT]
    fills = stat + dyn_short + dyn_long
    words[H_NFILL_STATIC] = len(stat)
    words[H_NFILL_SHORT_END] = len(stat) + len(dyn_short)
    offs = []
    for f in fills:
        offs.append(len(words))
        words.extend(f)
    words[H_NFILL] = len(offs)
    ptab = pool.table()
    words[H_NPOOL] = len(ptab) // 3
    words[H_POOLSIZE] = pool.size
    words.extend(ptab)  # pool table sits right before the fill-offset table
    words.extend(offs)
    words[H_WORDS] = len(words)
    program = np.asarray(words, dtype=np.int64).astype(np.int32)
    if len(program) > PASS_MAX_WORDS or pool.size > PASS_MAX_POOL:
        raise _ProgramTooLarge(f"pass program too large ({len(program)} words)")
    done = [gids[gi] for gi in done_local]
    step = PassStep(program=program, tile_bits=T, low_bits=L, gate_ids=[gates[gi].gid for gi in done],
                    n_subpasses=nsub, pool_elems=pool.size)  # fmt: skip
    return step, done
