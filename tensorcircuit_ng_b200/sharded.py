"""
Sharded statevector (BASELINE.json configs[3]): 2^n amplitudes over G = 2^g GPUs.

The reference has no sharded statevector (its multi-device examples shard Hamiltonian terms or
contraction slices, `examples/ng_whitepaper/VIA_sharding_vqe.py:34-37`,
`tensorcircuit/experimental.py:881-894`); the scheme is the north-star's: the g highest flat-index
bit positions are *global* (rank r holds the amplitudes whose global bits equal r).

  * diagonal gates and controls on global qubits are local work: every kernel takes `index_base`
    (= rank << n_local) and reads those bits from it;
  * a dense gate on a global qubit needs the qubit to be local first: an m-qubit global<->local
    swap is one exchange in which every rank keeps 2^-m of its shard and trades one sub-block with each
    of the 2^m - 1 ranks that differ in the swapped rank bits (`tcb_sv_pack_bits` -> P2P -> unpack,
    streamed through two staging buffers so packing overlaps the wire);
  * expectations are per-rank partial sums + one all-reduce.

The host scheduler cuts the gate stream into local segments (each compiled by passplan into fused
HBM passes under the *current* qubit layout) separated by swaps; which local qubits are evicted is
decided Belady-style (farthest next dense use).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import passplan
from .passplan import GateOp, _Frontier


@dataclass
class RunSegment:
    gate_ids: List[int]  # indices into the circuit's gate list, in a valid execution order
    pos_of: List[int]  # qubit -> flat bit position while this segment runs


@dataclass
class SwapSegment:
    pairs: List[Tuple[int, int]]  # (global position, local position) exchanged, one pair per swapped qubit
    pos_of_after: List[int]


@dataclass
class ShardedPlan:
    nqubits: int
    nglobal: int
    segments: List[Any]
    final_pos_of: List[int]
    cache: Dict[Any, Any] = field(default_factory=dict)  # compiled segments (device programs), per rank

    @property
    def n_swaps(self) -> int:
        return sum(1 for s in self.segments if isinstance(s, SwapSegment))

    @property
    def swapped_qubits(self) -> int:
        return sum(len(s.pairs) for s in self.segments if isinstance(s, SwapSegment))


def _needs_local(g: GateOp) -> Tuple[int, ...]:
    """Qubits of g that must be local for it to run (the rest may sit on rank bits)."""
    if g.is_diag:
        return ()
    if g.kind[0] == "ctrl" and g.kind[1] <= 2 and g.k == g.kind[1] + 1:
        return (g.qubits[-1],)
    return tuple(g.qubits)


def compile_sharded(gates: Sequence[GateOp], nqubits: int, nglobal: int, *, avoid_low: int = 1,
                    search_width: int = 4, search_budget: int = 2000) -> ShardedPlan:
    """Cut the gate stream into local segments and swaps.  `avoid_low`: the lowest local positions are
    never chosen for eviction (keeps the exchange 16-byte vectorised and the coalescing bits put).

    Which local qubits make room for the incoming ones is Belady's rule (farthest next dense use) — refined by
    a bounded depth-first search over the `search_width` next-best eviction sets at every swap when that saves
    a whole swap (31- and 32-qubit QAOA p = 8: 3 swaps and 4 segments instead of 4 and 5; every swap is an
    exchange of most of the shard plus the passes a short tail segment cannot fill)."""
    nl = nqubits - nglobal
    for gi, g in enumerate(gates):
        if g.gid < 0:
            g.gid = gi

    def runnable_for(pos_of: List[int]) -> Any:
        return lambda g: "local" if all(pos_of[q] < nl for q in _needs_local(g)) else None

    def swap_options(front: _Frontier, pos_of: List[int], everyone: bool = False) -> Tuple[List[int], List[int]]:
        """(incoming global qubits by next dense use, local eviction candidates by farthest next dense use).
        `everyone`: every global qubit with work left comes in, not only those needed sooner than the evicted."""
        def next_dense_use(q: int) -> int:
            qq = front.queues[q]
            for gi in qq[front.ptr[q]:]:
                if q in _needs_local(gates[gi]):
                    return gates[gi].gid
            return 1 << 60

        incoming = [q for q in range(nqubits) if pos_of[q] >= nl and next_dense_use(q) < (1 << 60)]
        incoming.sort(key=next_dense_use)
        pinned = set()  # qubits the blocked head gates need stay local
        for gi in front.heads():
            pinned.update(gates[gi].qubits)
        cand = [q for q in range(nqubits) if avoid_low <= pos_of[q] < nl and q not in pinned]
        cand.sort(key=lambda q: -next_dense_use(q))
        incoming = incoming[: len(cand)]
        if not incoming:
            raise RuntimeError("sharded planner stalled: a gate needs more local qubits than the shard has")
        # never evict a qubit that is needed sooner than the one it makes room for
        if everyone:
            return incoming, cand
        keep = [a for a, b in zip(incoming, cand) if next_dense_use(b) > next_dense_use(a)]
        return (keep or incoming[:1]), cand

    def build(choices: Optional[List[Tuple[Tuple[int, ...], Tuple[int, ...]]]]) -> ShardedPlan:
        """Greedy segments; swap k brings in / evicts `choices[k]` (default: the Belady sets)."""
        pos_of = [nqubits - 1 - q for q in range(nqubits)]
        front = _Frontier(gates, nqubits)
        segments: List[Any] = []
        k = 0
        while front.remaining > 0:
            run = front.simulate(set(), pos_of, 0, 1 << 60, commit=True, pred=runnable_for(pos_of))
            if run:
                segments.append(RunSegment([gi for gi, _ in run], list(pos_of)))
            if front.remaining == 0:
                break
            if choices is not None and k < len(choices):
                incoming, outgoing = list(choices[k][0]), list(choices[k][1])
            else:
                incoming, cand = swap_options(front, pos_of)
                outgoing = cand[: len(incoming)]
            k += 1
            pairs = []
            for qin, qout in zip(incoming, outgoing):
                pairs.append((pos_of[qin], pos_of[qout]))
                pos_of[qin], pos_of[qout] = pos_of[qout], pos_of[qin]
            segments.append(SwapSegment(pairs, list(pos_of)))
        return ShardedPlan(nqubits, nglobal, segments, list(pos_of))

    plan = build(None)
    if search_width <= 0 or plan.n_swaps <= 1:
        return plan
    # bounded DFS over eviction sets: fewer swaps than the greedy plan, or nothing
    best: List[Any] = [plan.n_swaps, None]
    budget = [search_budget]

    def rec(ptr: List[int], pos_of: List[int], chosen: List[Tuple[int, ...]]) -> None:
        front = _Frontier(gates, nqubits)
        front.ptr = list(ptr)
        front.remaining = len(gates)  # (only ptr matters for simulate / heads)
        front.simulate(set(), pos_of, 0, 1 << 60, commit=True, pred=runnable_for(pos_of))
        budget[0] -= 1
        if all(front.ptr[q] >= len(front.queues[q]) for q in range(nqubits)):
            if len(chosen) < best[0]:
                best[0], best[1] = len(chosen), list(chosen)
            return
        if len(chosen) + 1 >= best[0] or budget[0] <= 0:
            return
        import itertools

        seen_in = set()
        for everyone in (True, False):
            incoming, cand = swap_options(front, pos_of, everyone)
            if tuple(incoming) in seen_in:
                continue
            seen_in.add(tuple(incoming))
            m = len(incoming)
            for ti, outs in enumerate(itertools.combinations(cand[: m + search_width], m)):
                if ti >= 12 or budget[0] <= 0:
                    break
                p2 = list(pos_of)
                for qin, qout in zip(incoming, outs):
                    p2[qin], p2[qout] = p2[qout], p2[qin]
                rec(front.ptr, p2, chosen + [(tuple(incoming), tuple(outs))])

    rec([0] * nqubits, [nqubits - 1 - q for q in range(nqubits)], [])
    return build(best[1]) if best[1] is not None else plan


# ---------------------------------------------------------------------------------------------
class CudaExecutor:
    """Local work of one rank through the C ABI (the product path)."""

    def __init__(self, device: Any) -> None:
        import torch

        self.torch = torch
        self.device = device

    def zeros(self, n: int) -> Any:
        return self.torch.zeros(n, dtype=self.torch.complex64, device=self.device)

    def empty(self, n: int) -> Any:
        return self.torch.empty(n, dtype=self.torch.complex64, device=self.device)

    def set_one(self, state: Any) -> None:
        state[0] = 1.0

    def init_product(self, state: Any, nl: int, vecs: Any, nq: int, index_base: int) -> None:
        from . import _lib

        _lib.call("tcb_sv_init_product", state.data_ptr(), nl, vecs.data_ptr(), nq, index_base, _lib.stream_ptr())

    def compiled(self, ops: Sequence[GateOp], nq: int, nl: int, pos_of: Sequence[int], cache: Dict[Any, Any],
                 key: Any) -> Any:  # fmt: skip
        """The segment's fused-pass plan with its device programs (`svengine.CompiledCircuit`), cached."""
        from . import svengine

        cc = cache.get(key)
        if cc is None:
            plan = passplan.compile_plan(list(ops), nq, nbits_local=nl, pos_of=pos_of, **svengine.plan_options)
            if svengine.auto_low_bits and svengine.plan_options.get("low_bits") == 3 and nl >= 28:
                # same rule as svengine.compile_circuit: 16-byte segments when they save a pass of this segment
                n3 = sum(isinstance(st, passplan.PassStep) for st in plan.steps)
                if n3 >= 4:
                    alt = passplan.compile_plan(list(ops), nq, nbits_local=nl, pos_of=pos_of,
                                                **{**svengine.plan_options, "low_bits": 2})  # fmt: skip
                    n2 = sum(isinstance(st, passplan.PassStep) for st in alt.steps)
                    if len(alt.steps) - n2 <= len(plan.steps) - n3 and n2 * 1.04 < n3:
                        plan = alt
            cc = svengine.CompiledCircuit(plan, list(ops), self.device)
            cache[key] = cc
        return cc

    def run_gates(self, state: Any, ops: Sequence[GateOp], gatebuf: Any, nq: int, nl: int, pos_of: Sequence[int],
                  index_base: int, cache: Dict[Any, Any], key: Any) -> None:  # fmt: skip
        self.compiled(ops, nq, nl, pos_of, cache, key).run(state, gatebuf, index_base=index_base)

    def run_gates_generate(self, state: Any, ops: Sequence[GateOp], gatebuf: Any, nq: int, nl: int,
                           pos_of: Sequence[int], index_base: int, cache: Dict[Any, Any], key: Any, vecs: Any) -> bool:  # fmt: skip
        """The first segment of an evolution that starts from the product state `vecs` ([nq, 2] by flat bit
        position; None = |0...0>): its first pass generates the shard in shared memory (no init pass, no read).
        False when the segment does not open with a fused pass of the default tile size (caller initialises)."""
        from . import _lib, svengine

        cc = self.compiled(ops, nq, nl, pos_of, cache, key)
        steps = cc.plan.steps
        if not (svengine.fuse_start and steps and isinstance(steps[0], passplan.PassStep) and steps[0].tile_bits == 12):
            return False
        if vecs is None:
            vecs = self.torch.zeros(nq, 2, dtype=self.torch.complex64, device=self.device)
            vecs[:, 0] = 1.0
        st0 = steps[0]
        _lib.call("tcb_sv_run_pass_generate", state.data_ptr(), nl, cc.programs.data_ptr(), len(st0.program),
                  st0.tile_bits, st0.low_bits, st0.pool_elems, gatebuf.data_ptr(), index_base, vecs.data_ptr(), nq,
                  _lib.stream_ptr())  # fmt: skip
        cc.run_steps(state, gatebuf, index_base=index_base, first=1)
        return True

    def pack(self, state: Any, buf: Any, nl: int, sel: Sequence[int], pattern: int, first: int, count: int,
             unpack: bool) -> None:  # fmt: skip
        """`buf`: a tensor, or a raw device address (a peer GPU's staging buffer mapped into this process:
        the pack kernel then writes straight over NVLink)."""
        from . import _lib

        ptr = buf if isinstance(buf, int) else buf.data_ptr()
        _lib.call("tcb_sv_unpack_bits" if unpack else "tcb_sv_pack_bits", state.data_ptr(), ptr, nl,
                  len(sel), _lib.int_array(sel), pattern, first, count, _lib.stream_ptr())  # fmt: skip

    def expect_z(self, state: Any, nl: int, masks: Sequence[int], index_base: int) -> Any:
        from . import _lib

        t = self.torch
        zm = t.from_numpy(np.asarray(masks, dtype=np.int64)).to(self.device)
        out = t.zeros(len(masks), dtype=t.float64, device=self.device)
        _lib.call("tcb_sv_expect_z", state.data_ptr(), nl, 1, zm.data_ptr(), len(masks), index_base, out.data_ptr(),
                  _lib.stream_ptr())  # fmt: skip
        return out

    def norm2(self, state: Any) -> Any:
        return self.expect_z(state, int(state.numel()).bit_length() - 1, [0], 0)

    def read(self, state: Any, idx: int) -> Any:
        return state[idx : idx + 1].clone()


class TorchDistComm:
    """torch.distributed plumbing: NCCL on GPUs, gloo in the CPU test tier."""

    def __init__(self, group: Any = None) -> None:
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_async(self, sends: List[Tuple[int, Any]], recvs: List[Tuple[int, Any]]) -> Any:
        """Post the P2P batch (NCCL runs it on its own stream, after the work already queued on the
        current stream); `wait` makes the current stream wait for it."""
        import torch

        def real(t: Any) -> Any:  # complex tensors travel as (re, im) views of the same memory
            return torch.view_as_real(t) if t.is_complex() else t

        ops = []
        for peer, t in sends:
            ops.append(self.dist.P2POp(self.dist.isend, real(t), peer, self.group))
        for peer, t in recvs:
            ops.append(self.dist.P2POp(self.dist.irecv, real(t), peer, self.group))
        return self.dist.batch_isend_irecv(ops)

    def exchange(self, sends: List[Tuple[int, Any]], recvs: List[Tuple[int, Any]]) -> None:
        self.wait(self.exchange_async(sends, recvs))

    @staticmethod
    def wait(works: Any) -> None:
        for w in works or []:
            w.wait()

    def all_reduce_sum(self, t: Any) -> Any:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def peer_staging(self, nbytes: int, device: Any) -> Any:
        """A staging buffer of `nbytes` on every rank that all ranks can write directly (NVLink peer stores):
        torch symmetric memory (CUDA VMM allocations exchanged between the local processes).  Returns
        (local uint8 tensor, [device address of the buffer of rank r], barrier) or None when the ranks cannot
        map each other's memory (no NVLink / P2P, CPU test tier, TCB_SWAP_P2P=0); `barrier(channel)` is a
        device-side barrier of all ranks on the current stream."""
        import os

        import torch

        if os.environ.get("TCB_SWAP_P2P", "1") == "0" or getattr(device, "type", "cpu") != "cuda":
            return None
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm

            group = self.group if self.group is not None else self.dist.group.WORLD
            buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
            hdl = symm.rendezvous(buf, group=group.group_name)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        except Exception as exc:  # pylint: disable=broad-except
            import sys

            print(f"[tensorcircuit_ng_b200.sharded] peer staging unavailable ({type(exc).__name__}: {exc}); "
                  "qubit swaps go through NCCL send/recv", file=sys.stderr)  # fmt: skip
            ok = 0
        flag = torch.tensor([ok], device=device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if not int(flag[0]):
            return None
        return buf, ptrs, (lambda channel=0: hdl.barrier(channel=channel))


class ShardedStatevector:
    """One rank's shard of an n-qubit state + the collective operations on it."""

    def __init__(self, nqubits: int, comm: Any, executor: Any, chunk_elems: int = 1 << 26) -> None:
        world = comm.world
        if world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.n = nqubits
        self.g = world.bit_length() - 1
        self.nl = nqubits - self.g
        if self.nl < 1:
            raise ValueError("more ranks than amplitudes")
        self.comm = comm
        self.ex = executor
        self.rank = comm.rank
        self.index_base = self.rank << self.nl
        self.state = executor.zeros(1 << self.nl)
        if self.rank == 0:
            executor.set_one(self.state)
        self.pos_of = [nqubits - 1 - q for q in range(nqubits)]
        self.chunk = max(2, min(chunk_elems, 1 << self.nl))
        self._bufs: Optional[Tuple[Any, Any]] = None
        self._peer: Any = None  # peer-memory staging (buffer, addresses per rank, barrier); False = unavailable
        self._side: Any = None  # side stream of the peer-memory exchange (unpacking)
        self._pending: Any = None  # (product vectors,) of an initial state that has not been written yet (reset lazy)
        self._cache: Dict[Any, Any] = {}
        self.bytes_sent = 0
        self.swaps_done = 0

    def reset(self, vecs: Any = None, lazy: bool = False) -> None:
        """Back to |0...0> in the canonical layout (the shard buffer is reused), or to the product state
        prod_p vecs[p][x_p] (vecs: [n, 2] complex64 indexed by flat bit position; one write pass).
        `lazy`: do not write the shard now — the first segment of the next `run` generates it inside its first
        pass when it can (`CudaExecutor.run_gates_generate`); anything else that needs the shard writes it first."""
        self.pos_of = [self.n - 1 - q for q in range(self.n)]
        self._pending = None
        if lazy and hasattr(self.ex, "run_gates_generate"):
            self._pending = (vecs,)
            return
        self._write_initial(vecs)

    def _write_initial(self, vecs: Any) -> None:
        if vecs is not None:
            self.ex.init_product(self.state, self.nl, vecs, self.n, self.index_base)
            return
        self.state.zero_()
        if self.rank == 0:
            self.ex.set_one(self.state)

    def _materialize(self) -> None:
        if getattr(self, "_pending", None) is not None:
            (vecs,) = self._pending
            self._pending = None
            self._write_initial(vecs)

    # -- evolution ---------------------------------------------------------------------------
    def run(self, plan: ShardedPlan, gates: Sequence[GateOp], gatebuf: Any) -> None:
        assert plan.nqubits == self.n and plan.nglobal == self.g
        for si, seg in enumerate(plan.segments):
            if isinstance(seg, RunSegment):
                assert seg.pos_of == self.pos_of, "segment compiled for a different qubit layout"
                ops = [gates[gi] for gi in seg.gate_ids]
                if getattr(self, "_pending", None) is not None:
                    (vecs,) = self._pending
                    if self.ex.run_gates_generate(self.state, ops, gatebuf, self.n, self.nl, seg.pos_of,
                                                  self.index_base, plan.cache, (self.rank, si), vecs):  # fmt: skip
                        self._pending = None
                        continue
                    self._materialize()
                self.ex.run_gates(self.state, ops, gatebuf, self.n, self.nl, seg.pos_of, self.index_base,
                                  plan.cache, (self.rank, si))  # fmt: skip
            else:
                self._materialize()
                self.swap(seg.pairs)
                assert self.pos_of == seg.pos_of_after
        self._materialize()  # (a plan without segments: the shard is the initial state itself)

    def swap(self, pairs: Sequence[Tuple[int, int]]) -> None:
        """Exchange the qubits at global positions P_i with those at local positions p_i."""
        m = len(pairs)
        order = sorted(range(m), key=lambda i: pairs[i][1])  # pack kernel wants ascending local bits
        sel = [pairs[i][1] for i in order]
        jbits = [pairs[i][0] - self.nl for i in order]  # rank bit of each swapped global position
        assert all(0 <= j < self.g for j in jbits) and all(0 <= p < self.nl for p in sel)
        mine = 0
        for k, j in enumerate(jbits):
            mine |= ((self.rank >> j) & 1) << k
        block = 1 << (self.nl - m)
        chunk = min(self.chunk, block)
        nb = chunk * ((1 << m) - 1)
        if self._peer_swap(pairs, sel, jbits, mine, block, chunk, nb):
            return
        if self._bufs is None or self._bufs[0].numel() < nb:
            # two send + two receive staging buffers: chunk c+1 is packed while chunk c is on the wire
            self._bufs = tuple(self.ex.empty(nb) for _ in range(4))
        peers = []
        for x in range(1 << m):
            if x == mine:
                continue
            peer = self.rank
            for k, j in enumerate(jbits):
                peer = (peer & ~(1 << j)) | (((x >> k) & 1) << j)
            peers.append((x, peer))
        async_ok = hasattr(self.comm, "exchange_async")

        def post(ci: int, first: int, cnt: int) -> Any:
            sendbuf, recvbuf = self._bufs[ci & 1], self._bufs[2 + (ci & 1)]
            sends, recvs = [], []
            for i, (x, peer) in enumerate(peers):
                sb = sendbuf[i * chunk : i * chunk + cnt]
                # the amplitudes with local pattern x move to the rank whose bits are x ...
                self.ex.pack(self.state, sb, self.nl, sel, x, first, cnt, False)
                sends.append((peer, sb))
                recvs.append((peer, recvbuf[i * chunk : i * chunk + cnt]))
            if async_ok:
                return self.comm.exchange_async(sends, recvs)
            self.comm.exchange(sends, recvs)
            return None

        def finish(ci: int, first: int, cnt: int, works: Any) -> None:
            if async_ok:
                self.comm.wait(works)
            recvbuf = self._bufs[2 + (ci & 1)]
            for i, (x, peer) in enumerate(peers):
                # ... and that rank's block with local pattern `mine` takes their place
                self.ex.pack(self.state, recvbuf[i * chunk : i * chunk + cnt], self.nl, sel, x, first, cnt, True)
            self.bytes_sent += 8 * cnt * len(peers)

        pending = None
        for ci, first in enumerate(range(0, block, chunk)):
            cnt = min(chunk, block - first)
            works = post(ci, first, cnt)  # (reads positions this rank has not unpacked into yet)
            if pending is not None:
                finish(*pending)
            pending = (ci, first, cnt, works)
        if pending is not None:
            finish(*pending)
        self._finish_swap(pairs)

    def _finish_swap(self, pairs: Sequence[Tuple[int, int]]) -> None:
        inv = {p: q for q, p in enumerate(self.pos_of)}
        for P, p in pairs:
            qa, qb = inv[P], inv[p]
            self.pos_of[qa], self.pos_of[qb] = p, P
        self.swaps_done += 1

    def _peer_swap(self, pairs: Sequence[Tuple[int, int]], sel: List[int], jbits: List[int], mine: int, block: int,
                   chunk: int, nb: int) -> bool:  # fmt: skip
        """The exchange over peer memory: the pack kernel of every outgoing sub-block writes straight into the
        receiver's staging buffer (one local read + NVLink stores, no send buffer and no copy engine), a
        device-side barrier, then the receiver unpacks.  Two staging halves alternate, so the barrier of chunk
        c + 1 also tells that every rank is done unpacking chunk c - 1.  False: not available (NCCL path)."""
        if self._peer is None and hasattr(self.comm, "peer_staging"):
            # one staging buffer per process and size (every rank asks in the same order: the rendezvous is collective)
            key = (self.comm.world, 2 * 8 * self._peer_elems())
            if key not in _peer_staging_cache:
                _peer_staging_cache[key] = self.comm.peer_staging(key[1], getattr(self.state, "device", None)) or False
            self._peer = _peer_staging_cache[key]
        if not self._peer or nb > self._peer_elems():
            return False
        buf, ptrs, barrier = self._peer
        import torch

        local = buf.view(torch.complex64)
        half = self._peer_elems()
        m = len(sel)
        peers = []
        for x in range(1 << m):
            if x == mine:
                continue
            peer = self.rank
            for k, j in enumerate(jbits):
                peer = (peer & ~(1 << j)) | (((x >> k) & 1) << j)
            peers.append((x, peer))
        # unpacking runs on a side stream, so that it overlaps the next chunk's pushes (the wire never waits):
        #   main : push(c) .. [wait: own unpack(c - 1) done] barrier(c) .. push(c + 1) ..
        #   side : [wait: barrier(c)] unpack(c)
        # passing barrier(c) therefore means that every rank has unpacked chunk c - 1, whose staging half the
        # pushes of chunk c + 1 overwrite.
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        prev_unpacked = None
        for ci, first in enumerate(range(0, block, chunk)):
            cnt = min(chunk, block - first)
            par = ci & 1
            # destinations in the order x = mine ^ d, d = 1, 2, ..: at every d the ranks pair up (a perfect
            # matching), so no receiver takes two senders at once (in pattern order every rank starts on the
            # same receiver and the exchange runs at 1 / (#peers) of the link rate)
            for x, peer in sorted(peers, key=lambda xp: xp[0] ^ mine):
                # my slot in the receiver's buffer: its senders are ordered by pattern, its own pattern (x) left out
                slot = mine - (1 if x < mine else 0)
                dst = ptrs[peer] + 8 * (par * half + slot * chunk)
                self.ex.pack(self.state, dst, self.nl, sel, x, first, cnt, False)
            if prev_unpacked is not None:
                main.wait_event(prev_unpacked)
            barrier(par)
            arrived = torch.cuda.Event()
            arrived.record(main)
            side.wait_event(arrived)
            with torch.cuda.stream(side):
                for i, (x, peer) in enumerate(peers):
                    src = local[par * half + i * chunk : par * half + i * chunk + cnt]
                    self.ex.pack(self.state, src, self.nl, sel, x, first, cnt, True)
                prev_unpacked = torch.cuda.Event()
                prev_unpacked.record(side)
            self.bytes_sent += 8 * cnt * len(peers)
        main.wait_stream(side)
        barrier(2)  # (the staging halves may be reused by the next swap)
        self._finish_swap(pairs)
        return True

    def _peer_elems(self) -> int:
        """Complex elements of one staging half: a chunk from each of the (world - 1) possible senders."""
        return self.chunk * max(1, self.comm.world - 1)

    # -- read-out ----------------------------------------------------------------------------
    def _mask(self, qubits: Sequence[int]) -> int:
        m = 0
        for q in qubits:
            m ^= 1 << self.pos_of[q]
        return m

    def z_expectations(self, terms: Sequence[Sequence[int]]) -> Any:
        """<Z_S> for every term (qubit lists), all-reduced over the ranks (float64)."""
        self._materialize()
        out = self.ex.expect_z(self.state, self.nl, [self._mask(t) for t in terms], self.index_base)
        return self.comm.all_reduce_sum(out)

    def norm2(self) -> Any:
        self._materialize()
        out = self.ex.expect_z(self.state, self.nl, [0], self.index_base)
        return self.comm.all_reduce_sum(out)

    def amplitude(self, bits: Sequence[int]) -> Any:
        """Amplitude of the computational basis state `bits` (qubit 0 first), on every rank."""
        phys = 0
        for q, b in enumerate(bits):
            phys |= (int(b) & 1) << self.pos_of[q]
        owner, local = phys >> self.nl, phys & ((1 << self.nl) - 1)
        self._materialize()
        v = self.ex.read(self.state, local if owner == self.rank else 0)
        if owner != self.rank:
            v = v * 0
        # complex all-reduce as two reals (gloo has no complex sum on every version)
        vr = self._view_real(v)
        self.comm.all_reduce_sum(vr)
        return v

    @staticmethod
    def _view_real(v: Any) -> Any:
        import torch

        return torch.view_as_real(v)


# ---------------------------------------------------------------------------------------------
_plan_cache: Dict[Any, Any] = {}
_peer_staging_cache: Dict[Any, Any] = {}


def product_vectors(prefix: Sequence[Sequence[GateOp]], gatebuf: Any, n: int) -> Any:
    """[n, 2] complex64, row = flat bit position: U_k ... U_1 |0> for every qubit's leading 1q gates."""
    import torch

    v = torch.zeros(n, 2, dtype=torch.complex64, device=gatebuf.device)
    v[:, 0] = 1.0
    for q, gl in enumerate(prefix):
        for g in gl:
            if g.kind[0] == "diagvec":
                m = torch.diag(gatebuf[g.mat_off : g.mat_off + 2])
            else:
                m = gatebuf[g.mat_off : g.mat_off + 4].reshape(2, 2)
            v[n - 1 - q] = m.to(torch.complex64) @ v[n - 1 - q]
    return v


def evolve(circuit: Any, comm: Any = None, executor: Any = None, chunk_elems: int = 1 << 26,
           reuse: Optional[ShardedStatevector] = None) -> ShardedStatevector:
    """Run a `Circuit` (built with the ordinary gate API; construction never touches amplitudes,
    tensorcircuit/basecircuit.py:183-371) as a statevector sharded over the ranks of `comm`."""
    import torch

    from . import svengine

    if comm is None:
        comm = TorchDistComm()
    nodes, d_edges = circuit._copy()
    n, init, gates = svengine.extract_gate_stream(nodes, d_edges, max_qubits=48)
    if init is not None:
        raise NotImplementedError("sharded evolution starts from |0...0>")
    if executor is None:
        # (deferred gates are looked at through their parameter: building 600 matrices one by one would cost
        # more host time than the evolution itself)
        probe = [g[0]._lazy.theta if hasattr(g[0], "pending") and g[0].pending() else g[0].tensor for g in gates]
        executor = CudaExecutor(svengine.pick_device(probe))
    structure = tuple((g[1], k, int(math.prod(g[0].shape))) for g, k in zip(gates, svengine.gate_kinds(gates)))
    key = (n, comm.world, structure)
    hit = _plan_cache.get(key)
    if hit is None:
        allops, off = [], 0
        for gi, (qubits, kind, numel) in enumerate(structure):
            allops.append(GateOp(tuple(qubits), tuple(kind), off, gi))
            off += numel
        # every qubit's leading 1q gates act on |0> alone: they become the initial product state
        prefix, ops = svengine.split_prefix(allops, n)
        plan = compile_sharded(ops, n, comm.world.bit_length() - 1)
        if len(_plan_cache) > 32:
            _plan_cache.clear()
        _plan_cache[key] = hit = (plan, ops, prefix)
    plan, ops, prefix = hit
    if isinstance(executor, CudaExecutor):
        gatebuf = svengine.assemble_gatebuf([g[0] for g in gates], executor.device)  # one batched build per gate family
    else:
        gatebuf = torch.cat([g[0].tensor.reshape(-1).to(torch.complex64) for g in gates])
    vecs = product_vectors(prefix, gatebuf, n) if any(prefix) else None
    if reuse is not None and reuse.n == n:
        sv = reuse
        sv.reset(vecs, lazy=True)
    else:
        sv = ShardedStatevector(n, comm, executor, chunk_elems=chunk_elems)
        sv.reset(vecs, lazy=True)
    sv.prefix = prefix  # type: ignore[attr-defined]
    sv.run(plan, ops, gatebuf)
    sv.plan, sv.ops, sv.gatebuf = plan, ops, gatebuf  # type: ignore[attr-defined]
    return sv
