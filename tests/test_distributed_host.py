"""Host logic of the sliced / distributed contraction path (CPU): slice partition, plan schema,
slice-id -> index map, planners.  SURVEY §8a R13-R16."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_partition_matches_reference_rule(built):
    from tensorcircuit_ng_b200.experimental import slice_partition

    # tensorcircuit/experimental.py:877-894: S = ceil(n/G), row-major ids, -1 padding
    p = slice_partition(10, 4)
    assert p.shape == (4, 3) and p.dtype == np.int32
    assert p.tolist() == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, -1, -1]]
    assert slice_partition(1, 8).tolist() == [[0]] + [[-1]] * 7
    assert slice_partition(16, 2).tolist() == [list(range(8)), list(range(8, 16))]


def test_slice_values_mixed_radix(built):
    from tensorcircuit_ng_b200 import planner

    sd = {"a": 2, "b": 2, "c": 2}
    seen = set()
    for s in range(8):
        v = planner.slice_values(s, ["a", "b", "c"], sd)
        assert v == {"a": (s >> 2) & 1, "b": (s >> 1) & 1, "c": s & 1}  # last index fastest
        seen.add(tuple(sorted(v.items())))
    assert len(seen) == 8


def _nodes_fn(params):
    import tensorcircuit_ng_b200 as tc

    c = tc.Circuit(4)
    c.rx(range(4), theta=params["x"])
    c.cnot([0, 1, 2], [1, 2, 3])
    c.ry(range(4), theta=params["y"])
    return c.expectation_before([tc.gates.z(), [-1]], reuse=False)


def test_tree_data_schema_and_slicing(built):
    """The plan interchange format of tensorcircuit/experimental.py:947-953 (+ our labels)."""
    from tensorcircuit_ng_b200 import planner
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    params = {"x": np.ones([4], dtype=np.float32), "y": 0.3 * np.ones([4], dtype=np.float32)}
    td = DistributedContractor._get_tree_data(_nodes_fn, params, {"slicing_reconf_opts": {"target_size": 2**2}})
    assert {"inputs", "output", "size_dict", "path", "sliced_inds"} <= set(td)
    assert td["output"] == ()
    st = planner.path_stats(td["inputs"], td["output"], td["size_dict"], td["path"], list(td["sliced_inds"]))
    assert st["size"] <= 2**2 and st["nslices"] == 2 ** len(td["sliced_inds"]) and st["nslices"] >= 2
    # a linear path consumes every tensor exactly once
    n = len(td["inputs"])
    live = n
    for i, j in td["path"]:
        assert 0 <= i < live and 0 <= j < live and i != j
        live -= 1
    assert live == 1


@pytest.mark.parametrize("shape", [(3, 3, 6), (4, 4, 8)])
def test_planners_on_lattice_circuit(built, shape):
    """Elimination / wire-sweep trees must not be worse than the pairwise greedy on lattice circuits."""
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import cons, planner

    rows, cols, depth = shape
    n = rows * cols
    rng = np.random.default_rng(0)
    c = tc.Circuit(n)
    for l in range(depth):
        for q in range(n):
            c.rx(q, theta=float(rng.uniform(0, 6)))
        for r in range(rows):
            for q in range(cols):
                i = r * cols + q
                if l % 2 == 0 and q + 1 < cols and (q + r + l // 2) % 2 == 0:
                    c.cz(i, i + 1)
                if l % 2 == 1 and r + 1 < rows and (q + r + l // 2) % 2 == 0:
                    c.cz(i, i + cols)
    nodes = c.amplitude_before("0" * n)
    (inp, out, sd), sn = cons.get_tn_info(nodes)
    groups = cons.wire_groups(inp, sn)
    assert len(set(groups.values())) == n
    inp2, out2, sd2, _ = cons.diagonal_to_hyperedges(inp, out, sd, sn, [x.tensor for x in sn])
    assert sum(len(t) for t in inp2) < sum(len(t) for t in inp)
    te = planner.search_elimination(inp2, out2, sd2, groups=groups)
    tg = planner.search(inp2, out2, sd2)
    fe = planner.path_stats(inp2, out2, sd2, te["path"])["flops"]
    fg = planner.path_stats(inp2, out2, sd2, tg["path"])["flops"]
    assert fe <= 4 * fg


def _einsum_run(steps, arrays, inputs, fixed):
    tens = {}
    for i, (t, modes) in enumerate(zip(arrays, inputs)):
        idx = tuple(fixed[m] if m in fixed else slice(None) for m in modes)
        tens[i] = t.numpy()[idx]
    o = None
    for a, b, ta, tb, keep, o in steps:
        syms = {m: k for k, m in enumerate(dict.fromkeys(list(ta) + list(tb)))}
        tens[o] = np.einsum(tens.pop(a), [syms[m] for m in ta], tens.pop(b), [syms[m] for m in tb], [syms[m] for m in keep])
    assert len(tens) == 1
    return tens[o]


@pytest.mark.parametrize("shape,target", [((4, 8), 30), ((4, 8), 8), ((5, 10), 12)])
def test_schedule_chain_fusion_is_exact(built, monkeypatch, shape, target):
    """tnengine.build_schedule re-associates chains of skinny absorptions (A.S1).S2 -> A.(S1.S2): the value of
    every slice must not change (numpy einsum executes both schedules), and big-tensor traffic must not grow."""
    sys.path.insert(0, ROOT)
    import bench
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import planner, tnengine
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    rows, depth = shape
    monkeypatch.setattr(tnengine, "_FUSE_BIG", 1 << 5)  # let the rule fire at test sizes
    monkeypatch.setattr(tnengine, "_stream_ok", lambda big, small, keep: True)
    nodes_fn = lambda _: bench.build_rcs(tc, rows, rows, depth).amplitude_before("0" * (rows * rows))  # noqa: E731
    inp, out, sd, tensors, groups = DistributedContractor._network(nodes_fn, None, True)
    td = planner.search_elimination(inp, out, sd, target_size=2**target, groups=groups)
    sl = list(td["sliced_inds"])
    s0 = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sorted(sl), fuse=False)
    s1 = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sorted(sl), fuse=True)
    w0 = sum(2 ** len(k) for *_, k, _ in s0 if len(k) >= 5)
    w1 = sum(2 ** len(k) for *_, k, _ in s1 if len(k) >= 5)
    assert w1 <= w0
    for sid in range(min(4, 2 ** len(sl))):
        fixed = planner.slice_values(sid, sl, sd)
        v0, v1 = _einsum_run(s0, tensors, td["inputs"], fixed), _einsum_run(s1, tensors, td["inputs"], fixed)
        assert abs(v0 - v1) <= 1e-7 * max(1.0, abs(v0))


def test_schedule_fusion_halves_the_traffic_of_an_elimination_plan(built):
    """Variable-elimination trees (the committed 6x6 depth-14 plan) are made of skinny absorptions: fusing
    chains of them halves the bytes moved.  (The 7x7 depth-20 plan is a site-block sweep since round 2 — its
    steps are GEMM shaped and have nothing to fuse.)"""
    import pickle

    sys.path.insert(0, ROOT)
    import bench
    from tensorcircuit_ng_b200 import tnengine

    td = pickle.load(open(bench.rcs_plan_path(6, 6, 14, 26), "rb"))
    sl = sorted(td["sliced_inds"])
    s0 = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sl, fuse=False)
    s1 = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sl, fuse=True)
    traffic = lambda st: sum(8.0 * (2 ** len(ta) + 2 ** len(tb) + 2 ** len(k)) for _, _, ta, tb, k, _ in st)  # noqa: E731
    assert traffic(s1) < 0.7 * traffic(s0)


def test_layout_planning_puts_the_contracted_modes_lowest(built):
    """build_schedule orders every intermediate for its consumer: the modes the consuming step contracts are the
    LAST of the producer's mode list (lowest address bits), in the same relative order in both operands."""
    import pickle

    sys.path.insert(0, ROOT)
    import bench
    from tensorcircuit_ng_b200 import tnengine

    td = pickle.load(open(bench.rcs_plan_path(7, 7, 20, 30), "rb"))
    steps = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sorted(td["sliced_inds"]))
    nleaf = len(td["inputs"])
    big = 0
    for a, b, ta, tb, keep, o in steps:
        if a < nleaf or b < nleaf:
            continue  # leaves keep the layout they were given
        ks = [m for m in ta if m in set(tb) and m not in set(keep)]
        if not ks:
            continue
        assert ta[-len(ks):] == ks and tb[-len(ks):] == ks, (a, b)
        big += len(ta) >= 20
    assert big >= 20  # the boundary x site steps of the sweep

def test_numpy_tree_executor_matches_oracle_statevector():
    """oracle/tc_oracle/treeexec.py (the CPU leg of the contraction bench): sliced pairwise execution of a
    planner tree_data == the oracle statevector's amplitude, hyper-indices and slicing included."""
    import bench
    import tc_oracle as otc
    import tensorcircuit_ng_b200 as tc
    from tc_oracle.treeexec import contract_tree_numpy
    from tensorcircuit_ng_b200 import planner
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    rows, cols, depth = 3, 4, 8
    bits = "010011010110"
    nodes_fn = lambda _: bench.build_rcs(tc, rows, cols, depth).amplitude_before(bits)  # noqa: E731
    inp, out, sd, tensors, groups = DistributedContractor._network(nodes_fn, None, True)
    td = planner.search_elimination(inp, out, sd, target_size=2**4, groups=groups)
    assert len(td["sliced_inds"]) >= 1
    arrs = [t.detach().cpu().numpy() for t in tensors]
    nsl = int(np.prod([sd[x] for x in td["sliced_inds"]]))
    val = 0.0
    for s in range(nsl):
        fixed = planner.slice_values(s, list(td["sliced_inds"]), td["size_dict"])
        val = val + contract_tree_numpy(arrs, td["inputs"], td["output"], td["path"], fixed=fixed)
    ref = bench.build_rcs(otc, rows, cols, depth).wavefunction()[int(bits, 2)]
    assert abs(complex(val) - complex(ref)) < 1e-6
