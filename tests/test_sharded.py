"""Sharded statevector (SURVEY §8e, configs[3]): scheduler + swap logic on CPU.

The local work runs on the kernel-logic emulator (tests/sharded_emu.py, test infrastructure); the
exchange is checked both with an in-process thread world (G = 2, 4, 8) and with a real world_size-2
`gloo` process group.  The oracle is the single-process numpy restatement of the reference path."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build, random_layers  # noqa: E402


def _ops_and_buf(n, ops):
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import passplan, svengine

    c = build(tc, n, ops)
    nodes, d_edges = c._copy()
    nq, init, gates = svengine.extract_gate_stream(nodes, d_edges)
    gops, off = [], 0
    for gi, g in enumerate(gates):
        gops.append(passplan.GateOp(tuple(g[1]), tuple(svengine.gate_kind(g[0], g[2])), off, gi))
        off += int(g[0].tensor.numel())
    buf = np.concatenate([g[0].tensor.reshape(-1).numpy() for g in gates]).astype(np.complex64)
    return gops, torch.from_numpy(buf)


def _oracle_state(n, ops):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tc_oracle

    return build(tc_oracle, n, ops).wavefunction()


@pytest.mark.parametrize("n,world,seed", [(12, 2, 0), (13, 4, 1), (14, 8, 2), (12, 4, 3)])
def test_sharded_threads_match_oracle(built, n, world, seed):
    from sharded_emu import EmuExecutor, ThreadWorld, gather_logical
    from tensorcircuit_ng_b200 import sharded

    ops = random_layers(n, 3, seed)
    gops, buf = _ops_and_buf(n, ops)
    g = world.bit_length() - 1
    plan = sharded.compile_sharded(gops, n, g)
    assert plan.n_swaps >= 1  # random_layers touches every qubit with dense gates
    tw = ThreadWorld(world)
    shards = [None] * world
    results = [None] * world
    errors = []

    def worker(r):
        try:
            sv = sharded.ShardedStatevector(n, tw.comm(r), EmuExecutor(), chunk_elems=1 << 6)
            sv.run(plan, gops, buf)
            zz = sv.z_expectations([[0, n - 1], [1], [0, 1, 2]])
            amp = sv.amplitude([1, 0] * (n // 2) + [1] * (n % 2))
            shards[r] = sv
            results[r] = (zz.numpy().copy(), amp.numpy().copy(), float(sv.norm2()[0]))
        except Exception as e:  # pragma: no cover
            errors.append(e)
            raise

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors
    ref = _oracle_state(n, ops)
    full = gather_logical([s.state.numpy() for s in shards], shards[0].pos_of, n, n - g)
    assert np.abs(full - ref).max() <= 1e-5
    p = np.abs(ref.astype(np.complex128)) ** 2
    idx = np.arange(2**n)

    def zexp(qs):
        s = np.ones(2**n)
        for q in qs:
            s *= 1 - 2 * ((idx >> (n - 1 - q)) & 1)
        return float(np.sum(p * s))

    want = [zexp([0, n - 1]), zexp([1]), zexp([0, 1, 2])]
    bits = [1, 0] * (n // 2) + [1] * (n % 2)
    ai = int("".join(str(b) for b in bits), 2)
    for zz, amp, nrm in results:
        assert np.allclose(zz, want, atol=1e-6)
        assert abs(amp[0] - ref[ai]) <= 1e-5
        assert abs(nrm - 1.0) <= 1e-5


def test_sharded_plan_properties(built):
    from tensorcircuit_ng_b200 import sharded
    from tensorcircuit_ng_b200.passplan import GateOp

    n, g = 10, 2
    # diagonal gates and controls on global qubits never force a swap
    gops = [GateOp((0, 5), ("diag",), 0), GateOp((1, 6), ("ctrl", 1, 1), 16), GateOp((0,), ("diag",), 32)]
    plan = sharded.compile_sharded(gops, n, g)
    assert plan.n_swaps == 0 and len(plan.segments) == 1
    # a dense gate on a global qubit does, and the evicted qubit is one that is not needed again
    gops = [GateOp((0,), ("dense",), 0), GateOp((9,), ("dense",), 4), GateOp((5,), ("dense",), 8)]
    plan = sharded.compile_sharded(gops, n, g)
    assert plan.n_swaps == 1
    sw = [s for s in plan.segments if isinstance(s, sharded.SwapSegment)][0]
    assert all(P >= n - g and 1 <= p < n - g for P, p in sw.pairs)
    evicted = [q for q in range(n) if plan.final_pos_of[q] >= n - g]
    assert 9 not in evicted and 5 not in evicted and 0 not in evicted


def test_sharded_eviction_search_saves_a_swap(built):
    """31-qubit QAOA p = 8 (the N = 2 scaling workload): Belady eviction alone needs 4 swaps and leaves a one-gate
    tail segment; the bounded search over eviction sets finds a 3-swap schedule.  Every segment stays a valid
    execution order (all gates once, dense gates only on local qubits)."""
    import networkx as nx

    from tensorcircuit_ng_b200 import sharded
    from tensorcircuit_ng_b200.passplan import GateOp

    n, g, p = 31, 1, 8
    gr = nx.random_regular_graph(3, n - 1, seed=0)
    gr.add_edges_from([(n - 1, 0), (n - 1, 1), (n - 1, 2)])
    gops, off = [], 0
    for _ in range(p):
        for a, b in gr.edges:
            gops.append(GateOp((int(a), int(b)), ("diag",), off))
            off += 16
        for q in range(n):
            gops.append(GateOp((q,), ("dense",), off))
            off += 4
    greedy = sharded.compile_sharded(gops, n, g, search_width=0)
    plan = sharded.compile_sharded(gops, n, g)
    assert greedy.n_swaps == 4 and plan.n_swaps == 3
    seen = []
    for seg in plan.segments:
        if isinstance(seg, sharded.RunSegment):
            for gi in seg.gate_ids:
                if not gops[gi].is_diag:
                    assert all(seg.pos_of[q] < n - g for q in gops[gi].qubits)
            seen += seg.gate_ids
    assert sorted(seen) == list(range(len(gops)))
    done = [0] * n  # per-qubit program order is kept
    queues = [[gi for gi, gt in enumerate(gops) if q in gt.qubits] for q in range(n)]
    for gi in seen:
        for q in gops[gi].qubits:
            assert queues[q][done[q]] == gi
            done[q] += 1


def _gloo_worker(rank, world, port, n, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from sharded_emu import EmuExecutor
    from tensorcircuit_ng_b200 import sharded

    ops = random_layers(n, 2, seed)
    gops, buf = _ops_and_buf(n, ops)
    plan = sharded.compile_sharded(gops, n, 1)
    sv = sharded.ShardedStatevector(n, sharded.TorchDistComm(), EmuExecutor(), chunk_elems=1 << 7)
    sv.run(plan, gops, buf)
    zz = sv.z_expectations([[0, n - 1], [2]])
    torch.save({"state": sv.state, "pos_of": sv.pos_of, "zz": zz, "swaps": sv.swaps_done}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_sharded_gloo_world2(built, tmp_path):
    import torch.multiprocessing as mp
    from sharded_emu import gather_logical

    n, seed, world = 11, 7, 2
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "shard")
    mp.spawn(_gloo_worker, args=(world, port, n, seed, out), nprocs=world, join=True)
    parts = [torch.load(f"{out}.{r}", weights_only=False) for r in range(world)]
    assert parts[0]["swaps"] >= 1
    ref = _oracle_state(n, random_layers(n, 2, seed))
    full = gather_logical([p["state"].numpy() for p in parts], parts[0]["pos_of"], n, n - 1)
    assert np.abs(full - ref).max() <= 1e-5
    assert torch.allclose(parts[0]["zz"], parts[1]["zz"])


def test_sharded_evolve_api_with_product_start(built):
    """`sharded.evolve(circuit, ...)`: the Circuit front end + leading 1q gates folded into the initial
    product state of every shard (rank bits read from index_base)."""
    import tensorcircuit_ng_b200 as tc
    from sharded_emu import EmuExecutor, ThreadWorld, gather_logical
    from tensorcircuit_ng_b200 import sharded

    n, world = 12, 4
    ops = [("h", [q], {}) for q in range(n)] + [("rz", [0], {"theta": 0.3})] + random_layers(n, 2, 9)
    tw = ThreadWorld(world)
    shards = [None] * world
    errors = []

    def worker(r):
        try:
            c = build(tc, n, ops)
            shards[r] = sharded.evolve(c, tw.comm(r), EmuExecutor(), chunk_elems=1 << 6)
        except Exception as e:  # pragma: no cover
            errors.append(e)
            raise

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors
    assert sum(len(p) for p in shards[0].prefix) >= n  # the h layer (+ rz) was absorbed
    ref = _oracle_state(n, ops)
    full = gather_logical([s.state.numpy() for s in shards], shards[0].pos_of, n, n - 2)
    assert np.abs(full - ref).max() <= 1e-5
