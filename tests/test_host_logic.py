"""CPU tier: the C-ABI library loads and exports every declared symbol; the pass planner's
programs, executed by the kernel-logic emulator (the SAME pass_core.cuh the GPU compiles),
reproduce the oracle; gate-stream extraction and contractor error behaviour match the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import tc_oracle
from helpers import brickwork, build, oracle_state, qaoa, random_layers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from tensorcircuit_ng_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "tcb200.h")).read()
    declared = set(re.findall(r"\b(tcb_[a-z0-9_]+)\s*\(", header))
    declared.discard("tcb_contract_desc")
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tcb200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"
    assert lib.tcb_abi_version() == 1


def test_no_cpu_fallback(built):
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200._lib import EngineError

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    c = tc.Circuit(3)
    c.h(0)
    with pytest.raises(EngineError, match="no CPU fallback"):
        c.wavefunction()


def _emu():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libpass_emu.so"))
    return lib


def _product_stream(n, ops):
    """Build the product circuit on CPU tensors and extract (structure, gate buffer)."""
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import svengine

    c = build(tc, n, ops)
    nodes, d_edges = c._copy()
    nq, init, gates = svengine.extract_gate_stream(nodes, d_edges)
    assert nq == n and init is None
    structure = [(g[1], svengine.gate_kind(g[0], g[2]), int(g[0].tensor.numel())) for g in gates]
    buf = np.concatenate([g[0].tensor.reshape(-1).numpy() for g in gates]).astype(np.complex64)
    return structure, buf


def _run_emulated(n, structure, buf, **opts):
    from tensorcircuit_ng_b200 import passplan

    gops, off = [], 0
    for gi, (qubits, kind, numel) in enumerate(structure):
        gops.append(passplan.GateOp(tuple(qubits), tuple(kind), off, gi))
        off += numel
    plan = passplan.compile_plan(gops, n, **opts)
    state = np.zeros(2**n, dtype=np.complex64)
    state[0] = 1
    emu = _emu()
    for st in plan.steps:
        if isinstance(st, passplan.PassStep):
            prog = np.ascontiguousarray(st.program)
            rc = emu.emu_run_pass(
                state.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_longlong(1),
                prog.ctypes.data_as(ctypes.c_void_p), len(prog), st.tile_bits, st.low_bits,
                buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(0), ctypes.c_ulonglong(0),
            )  # fmt: skip
            assert rc == 0
        else:
            g = st.gate
            k = g.k
            if g.is_diag:
                d = buf[g.mat_off : g.mat_off + (2**k if g.kind[0] == "diagvec" else 4**k)]
                if g.kind[0] != "diagvec":
                    d = np.diag(d.reshape(2**k, 2**k))
                m = np.diag(d).reshape([2] * (2 * k))
            else:
                m = buf[g.mat_off : g.mat_off + 4**k].reshape([2] * (2 * k))
            psi = np.tensordot(m, state.reshape([2] * n), axes=[list(range(k, 2 * k)), list(g.qubits)])
            state = np.ascontiguousarray(np.moveaxis(psi, list(range(k)), list(g.qubits))).reshape(-1)
    return state, plan


@pytest.mark.parametrize("n,depth,seed", [(4, 3, 0), (9, 3, 1), (10, 3, 2), (12, 4, 3), (14, 3, 4), (16, 2, 5),
                                          (11, 5, 6), (12, 5, 7), (13, 4, 8), (10, 6, 9), (15, 3, 10), (12, 6, 11)])
def test_planner_and_kernel_logic_match_oracle(built, n, depth, seed):
    ops = random_layers(n, depth, seed)
    ref = oracle_state(n, ops)
    structure, buf = _product_stream(n, ops)
    out, plan = _run_emulated(n, structure, buf)
    assert plan.n_gates == len(ops)
    np.testing.assert_allclose(out, ref, atol=1e-5)


@pytest.mark.parametrize("tile_bits,low_bits", [(10, 4), (11, 3), (12, 5), (13, 4)])
def test_tile_geometries(built, tile_bits, low_bits):
    n = 15
    ops = random_layers(n, 2, 11)
    ref = oracle_state(n, ops)
    structure, buf = _product_stream(n, ops)
    out, plan = _run_emulated(n, structure, buf, tile_bits=tile_bits, low_bits=low_bits)
    np.testing.assert_allclose(out, ref, atol=1e-5)


def test_config1_brickwork_plan(built):
    """BASELINE.json config 1 (20-qubit brickwork, depth 10): plan fuses 295 gates into few passes."""
    n = 20
    ops = brickwork(n, 10)
    structure, buf = _product_stream(n, ops)
    out, plan = _run_emulated(n, structure, buf)
    ref = oracle_state(n, ops)
    np.testing.assert_allclose(out, ref, atol=1e-5)
    assert plan.n_gates == 295
    assert plan.n_passes <= 12


def test_qaoa_plan_fusion(built):
    """BASELINE.json config 3 at reduced width: every ZZ layer rides along a dense pass."""
    n, p = 16, 3
    ops, edges = qaoa(n, p)
    structure, buf = _product_stream(n, ops)
    assert all(kind == ("diag",) for (q, kind, _) in structure if len(q) == 2)
    out, plan = _run_emulated(n, structure, buf)
    ref = oracle_state(n, ops)
    np.testing.assert_allclose(out, ref, atol=1e-5)
    assert plan.n_passes <= 2 * p + 2


def test_plan_is_deterministic_and_cached(built):
    from tensorcircuit_ng_b200 import passplan

    n = 14
    structure, _ = _product_stream(n, random_layers(n, 3, 7))
    progs = []
    for _ in range(2):
        gops, off = [], 0
        for gi, (qubits, kind, numel) in enumerate(structure):
            gops.append(passplan.GateOp(tuple(qubits), tuple(kind), off, gi))
            off += numel
        plan = passplan.compile_plan(gops, n)
        progs.append([s.program.tobytes() for s in plan.steps if isinstance(s, passplan.PassStep)])
    assert progs[0] == progs[1]


def test_extract_rejects_non_circuit_networks(built):
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import svengine

    c = tc.Circuit(3)
    c.h(0)
    c.cnot(0, 1)
    nodes = c.amplitude_before("010")
    with pytest.raises(svengine.NotCircuitShaped):
        svengine.extract_gate_stream(nodes, [])
    nodes = c.expectation_before([tc.gates.z(), [0]], reuse=False)
    with pytest.raises(svengine.NotCircuitShaped):
        svengine.extract_gate_stream(nodes, [])


def test_contractor_edge_order_errors(built):
    """Same ValueErrors as the reference (tensorcircuit/cons.py:886-896) — raised before any GPU work."""
    import tensorcircuit_ng_b200 as tc

    c = tc.Circuit(2)
    c.h(0)
    nodes, d_edges = c._copy()
    with pytest.raises(ValueError, match="more than one remaining edge"):
        tc.cons.contractor(nodes)
    nodes, d_edges = c._copy()
    with pytest.raises(ValueError, match="output edges are not equal"):
        tc.cons.contractor(nodes, output_edge_order=d_edges[:1])
    with pytest.raises(ValueError, match="duplicate qubits"):
        c.cnot(1, 1)
    with pytest.raises(ValueError, match="Cannot measure two operators in one index"):
        c.expectation_before([tc.gates.z(), [0]], [tc.gates.x(), [0]], reuse=False)


def test_node_capture_counts(built):
    """tests/test_miscs.py:307-324 of the reference: 7 and 9 nodes."""
    import tensorcircuit_ng_b200 as tc

    with tc.cons.runtime_nodes_capture() as captured:
        c = tc.Circuit(3)
        c.h(0)
        c.amplitude("010")
    assert len(captured["nodes"]) == 7

    @tc.cons.function_nodes_capture
    def exp(theta):
        c = tc.Circuit(3)
        c.h(0)
        return c.expectation_ps(z=[-3], reuse=False)

    assert len(exp(0.3)) == 9


def test_contractor_scoping(built):
    """tests/test_backends.py:1342-1358 of the reference: with-level and function-level contractor scoping."""
    import tensorcircuit_ng_b200 as tc

    base = tc.cons.contractor
    with tc.runtime_contractor("plain") as cf:
        assert tc.cons.contractor is cf and tc.circuit.contractor is cf
    assert tc.cons.contractor is base and tc.circuit.contractor is base

    @tc.set_function_contractor("plain")
    def f():
        return tc.circuit.contractor

    assert f() is tc.cons.plain_contractor
    assert tc.circuit.contractor is base


def test_gate_tensors_match_oracle(built):
    """Every gate factory of the product equals the oracle's (reference gates.py) tensor."""
    import tensorcircuit_ng_b200 as tc
    from tc_oracle import gates as og

    pg = tc.gates
    for name in ["i", "x", "y", "z", "h", "s", "t", "sd", "td", "wroot", "cnot", "cz", "cy", "swap", "toffoli",
                 "fredkin", "ox", "oy", "oz"]:  # fmt: skip
        np.testing.assert_allclose(getattr(pg, name)().tensor.numpy(), getattr(og, name)().tensor, atol=1e-7, err_msg=name)
    th = 0.37
    for name in ["rx", "ry", "rz", "phase", "iswap", "rzz", "rxx", "ryy", "crx", "cry", "crz", "cphase", "orx", "ory", "orz"]:
        a = getattr(pg, name + "_gate")(theta=th).tensor.numpy()
        b = getattr(og, name + "_gate")(theta=th).tensor
        np.testing.assert_allclose(a, b, atol=1e-6, err_msg=name)
    np.testing.assert_allclose(pg.u_gate(theta=0.3, phi=0.2, lbd=-0.4).tensor.numpy(), og.u_gate(0.3, 0.2, -0.4).tensor, atol=1e-6)
    np.testing.assert_allclose(pg.r_gate(theta=0.3, alpha=0.2, phi=-0.4).tensor.numpy(), og.r_gate(0.3, 0.2, -0.4).tensor, atol=1e-6)
    np.testing.assert_allclose(pg.cr_gate(theta=0.3, alpha=0.2, phi=-0.4).tensor.numpy(), og.cr_gate(0.3, 0.2, -0.4).tensor, atol=1e-6)
    np.testing.assert_allclose(pg.cu_gate(theta=0.3, phi=0.2, lbd=-0.4).tensor.numpy(), og.cu_gate(theta=0.3, phi=0.2, lbd=-0.4).tensor, atol=1e-6)
    np.testing.assert_allclose(pg.exp1_gate(og._xx_matrix, 0.3).tensor.numpy(), og.exp1_gate(og._xx_matrix, 0.3).tensor, atol=1e-6)
    np.testing.assert_allclose(pg.exp_gate(np.diag([1.0, -1, -1, 1]), 0.3).tensor.numpy(), og.exp_gate(np.diag([1.0, -1, -1, 1]), 0.3).tensor, atol=5e-6)  # complex64 matrix_exp


def test_gate_kind_hints_are_sound(built):
    """A `diag` / `ctrl` hint must be structurally true for the tensor it is attached to."""
    import tensorcircuit_ng_b200 as tc

    pg = tc.gates
    th = 1.234
    gates = [getattr(pg, n)() for n in ["i", "x", "y", "z", "h", "s", "t", "sd", "td", "cnot", "cz", "cy", "swap",
                                         "toffoli", "fredkin", "ox", "oy", "oz"]]  # fmt: skip
    gates += [getattr(pg, n + "_gate")(theta=th) for n in ["rx", "ry", "rz", "phase", "rzz", "rxx", "crx", "cry", "crz",
                                                            "cphase", "orx", "ory", "orz", "iswap"]]  # fmt: skip
    gates += [pg.cr_gate(theta=th, alpha=0.3, phi=0.2), pg.cu_gate(theta=th, phi=0.3, lbd=0.1),
              pg.exp1_gate(pg._zz_matrix, th), pg.exp1_gate(pg._xx_matrix, th), pg.any_gate(np.diag([1, 1j])),
              pg.any_gate(np.array([[0, 1], [1, 0]]))]  # fmt: skip
    for g in gates:
        kind = g._b200_kind
        d = int(round(np.sqrt(g.tensor.numel())))
        m = g.tensor.reshape(d, d).numpy()
        if kind[0] == "diag":
            assert np.allclose(m, np.diag(np.diag(m))), g.name
        elif kind[0] == "ctrl":
            nctrl, pol = kind[1], kind[2]
            polval = 0
            for i in range(nctrl):
                polval = (polval << 1) | ((pol >> i) & 1)
            rest = m.copy()
            rest[2 * polval : 2 * polval + 2, 2 * polval : 2 * polval + 2] = np.eye(2)
            assert np.allclose(rest, np.eye(d)), g.name


def test_planner_greedy_path_is_valid_and_competitive(built):
    from tc_oracle import cons as ocons
    from tc_oracle import paths
    from tensorcircuit_ng_b200 import planner

    n, d = 10, 4
    ops = [("h", [i], {}) for i in range(n)]
    from tc_oracle import gates as og

    for j in range(d):
        ops += [("exp1", [i, i + 1], {"unitary": og._zz_matrix, "theta": 1.0}) for i in range(n - 1)]
        ops += [("rx", [i], {"theta": 1.0}) for i in range(n)]
    c = build(tc_oracle, n, ops)
    nodes, _ = c._copy()
    (inp, out, sd), sorted_nodes = ocons.get_tn_info(nodes)
    path = planner.greedy(inp, out, sd)
    res = ocons.contract_path_einsum([x.tensor for x in sorted_nodes], ["".join(i) for i in inp], "".join(out), path)
    np.testing.assert_allclose(res.reshape(-1), c.wavefunction(), atol=1e-5)
    ours = planner.path_stats(inp, out, sd, path)
    ref = paths.path_cost(inp, out, sd, paths.greedy(inp, out, sd))
    assert ours["flops"] <= 1.5 * ref["flops"]
    # sliced search honours the size bound on a scalar network
    c2 = build(tc_oracle, n, ops)
    nodes = c2.amplitude_before("0" * n)
    (inp, out, sd), _ = ocons.get_tn_info(nodes)
    td = planner.search(inp, out, sd, target_size=2**4)
    st = planner.path_stats(inp, out, sd, td["path"], list(td["sliced_inds"]))
    assert st["size"] <= 2**4 and len(td["sliced_inds"]) >= 1
    vals = planner.slice_values(5, list(td["sliced_inds"]), sd)
    assert set(vals) == set(td["sliced_inds"])


def test_lazy_parametrised_gates_match_eager_values_and_gradients():
    """gates.LazyGate + svengine.assemble_gatebuf (batched per-family construction) against the eager
    per-gate factories: same gate buffer, same parameter gradients, same node behaviour."""
    import torch

    from tensorcircuit_ng_b200 import gates, svengine
    import tensorcircuit_ng_b200 as tc

    torch.manual_seed(3)
    n = 5
    p0 = torch.randn(3, 2, n, dtype=torch.float32)
    zz = np.kron(np.array([[1.0, 0], [0, -1.0]]), np.array([[1.0, 0], [0, -1.0]]))

    def build(p):
        c = tc.Circuit(n)
        for q in range(n):
            c.h(q)
        for l in range(3):
            for q in range(n - 1):
                c.rzz(q, q + 1, theta=p[l, 0, q])
            for q in range(n):
                c.rx(q, theta=p[l, 1, q])
            c.ry(0, theta=p[l, 0, n - 1])
            c.rz(1, theta=p[l, 1, 0] * 2.0)
            c.exp1(2, 3, unitary=zz, theta=p[l, 1, 1])
            c.cnot(0, 1)
        return [g for g in c._nodes if not g.name.startswith("qb-")]

    def run(lazy):
        old = gates.lazy_parametrised
        gates.lazy_parametrised = lazy
        try:
            p = p0.clone().requires_grad_(True)
            nodes = build(p)
            if lazy:
                assert sum(isinstance(g, gates.LazyGate) and g.pending() for g in nodes) == 3 * (2 * n - 1 + 3)
                assert nodes[n].shape == (2, 2, 2, 2) and nodes[n].get_rank() == 4 and nodes[n]._b200_kind == ("diag",)
                cp = nodes[n].copy()
                assert cp.pending() and cp._stable_id_ != nodes[n]._stable_id_
            buf = svengine.assemble_gatebuf(nodes, torch.device("cpu"))
            w = torch.arange(buf.numel(), dtype=torch.float32) * 0.01
            loss = (buf.real * w).sum() + (buf.imag * w.flip(0)).sum()
            (g,) = torch.autograd.grad(loss, p)
            return buf.detach(), g, nodes
        finally:
            gates.lazy_parametrised = old

    b1, g1, nodes1 = run(True)
    b0, g0, _ = run(False)
    assert b1.shape == b0.shape
    assert float((b1 - b0).abs().max()) < 1e-6
    assert float((g1 - g0).abs().max()) < 1e-5
    # a single access materialises one gate with the same values; conjugated copies are ordinary gates
    lz = next(g for g in nodes1 if isinstance(g, gates.LazyGate))
    t = lz.tensor
    assert not lz.pending() and tuple(t.shape) == lz.shape
    cj = lz.copy(conjugate=True)
    assert not isinstance(cj, gates.LazyGate) and torch.equal(cj.tensor.resolve_conj(), t.conj().resolve_conj())


def test_diagonal_only_subcircuit_plans(built):
    """The backward walk un-applies a run of diagonal gates as a circuit of its own (autograd._DiagRun): a plan
    made of diagonal gates only, applied to an arbitrary state, must equal the elementwise product."""
    from tensorcircuit_ng_b200 import passplan

    emu = _emu()
    for n, seed in [(6, 0), (13, 1), (15, 2)]:
        rng = np.random.default_rng(seed)
        gops, mats, off = [], [], 0
        pairs = [(q, q + 1) for q in range(n - 1)] + [(int(rng.integers(0, n)),) for _ in range(4)] + [(n - 1, 0)]
        for gi, qs in enumerate(pairs):
            d = np.exp(1j * rng.uniform(0, 2 * np.pi, size=2 ** len(qs))).astype(np.complex64)
            mats.append(np.diag(d).astype(np.complex64).reshape(-1))
            gops.append(passplan.GateOp(tuple(qs), ("diag",), off, gi))
            off += 4 ** len(qs)
        buf = np.concatenate(mats)
        plan = passplan.compile_plan(gops, n)
        state = (rng.normal(size=2**n) + 1j * rng.normal(size=2**n)).astype(np.complex64)
        want = state.astype(np.complex128).reshape([2] * n)
        for g, m in zip(gops, mats):
            k = g.k
            dd = np.diag(m.reshape(2**k, 2**k)).reshape([2] * k)
            shape = [1] * n
            for ax, q in enumerate(g.qubits):
                shape[q] = 2
            want = want * np.transpose(dd, np.argsort(np.argsort(g.qubits))).reshape(shape) if k > 1 else want * dd.reshape(shape)
        npass = 0
        for st in plan.steps:
            if not isinstance(st, passplan.PassStep):  # (tiny states: one launch per gate)
                g = st.gate
                dd = np.diag(buf[g.mat_off : g.mat_off + 4**g.k].reshape(2**g.k, 2**g.k)).reshape([2] * g.k)
                shape = [1] * n
                for q in g.qubits:
                    shape[q] = 2
                dd = np.transpose(dd, np.argsort(np.argsort(g.qubits))) if g.k > 1 else dd
                state = (state.reshape([2] * n) * dd.reshape(shape)).reshape(-1).astype(np.complex64)
                continue
            npass += 1
            prog = np.ascontiguousarray(st.program)
            rc = emu.emu_run_pass(state.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_longlong(1),
                                  prog.ctypes.data_as(ctypes.c_void_p), len(prog), st.tile_bits, st.low_bits,
                                  buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(0), ctypes.c_ulonglong(0))  # fmt: skip
            assert rc == 0
        assert n < 12 or 1 <= npass <= 3
        assert np.abs(state - want.reshape(-1)).max() < 2e-5 * np.abs(want).max()


def test_second_plan_with_16_byte_segments_is_kept_only_when_it_saves_a_pass(built):
    """svengine.compile_circuit plans states of >= 28 qubits a second time with low_bits = 2 and keeps that plan
    only when it has fewer passes (30-qubit QAOA p=8: 18 instead of 19); smaller states keep low_bits = 3."""
    import bench
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import passplan, svengine

    if not svengine.auto_low_bits or svengine.plan_options.get("low_bits") != 3:
        pytest.skip("TCB_LOW_BITS is set")
    zz = np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0]))
    for n, p, want_low in [(30, 8, 2), (24, 4, 3)]:
        edges, gam, bet = bench.qaoa_problem(n, p)
        c = bench.build_qaoa(tc, n, edges, torch.from_numpy(gam), torch.from_numpy(bet), zz)
        nodes, e = c._copy()
        nn, init, gates = svengine.extract_gate_stream(nodes, e)
        structure = [(g[1], svengine.gate_kind(g[0], g[2]), int(np.prod(g[0].shape))) for g in gates]
        cc = svengine.compile_circuit(nn, structure, torch.device("cpu"), absorb_prefix=True)
        passes = [s for s in cc.plan.steps if isinstance(s, passplan.PassStep)]
        assert {s.low_bits for s in passes} == {want_low}
        if n == 30:
            _, rest = svengine.split_prefix(cc.ops, nn)
            base = passplan.compile_plan(rest, nn, **svengine.plan_options)
            n3 = sum(isinstance(s, passplan.PassStep) for s in base.steps)
            assert len(passes) < n3


def test_adjoint_segments_partition_the_circuit(built):
    """autograd._AdjointTables: every gate lands in exactly one segment; diagonal runs hold only diagonal gates,
    one-qubit runs distinct qubits, constant runs only constants; a one-qubit run directly followed by a diagonal
    run gets a fused un-apply plan."""
    import torch

    from tensorcircuit_ng_b200 import autograd, svengine

    n = 14
    structure, const = [], []

    def add(qubits, kind, is_const):
        structure.append((tuple(qubits), (kind,), 4 ** len(qubits)))
        const.append(is_const)

    for q in range(n):
        add([q], "dense", True)  # H layer
    for l in range(2):
        for q in range(n - 1):
            add([q, q + 1], "diag", False)  # rzz layer
        for q in range(n):
            add([q], "dense", False)  # rx layer
        for q in range(n - 1):
            add([q, q + 1], "dense", True)  # CNOT ladder
        add([0, 1, 2], "dense", True)  # toffoli
        for q in range(n):  # ry(q) rz(q): qubits repeat
            add([q], "dense", False)
            add([q], "diag", False)
        add([3, 4], "dense", False)  # a trainable two-qubit gate: gate-by-gate
    cc = svengine.compile_circuit(n, structure, torch.device("cpu"), absorb_prefix=False)
    tabs = autograd._AdjointTables(cc, sum(s[2] for s in structure), torch.device("cpu"), tuple(const))
    seen = []
    kinds = set()
    for seg in tabs.segments:
        if isinstance(seg, tuple):
            idx = list(range(seg[1], seg[2]))
            kinds.add("G")
        elif isinstance(seg, autograd._OneQubitRun) and seg.first == -1:
            idx = None  # rounds: indices live in seg.dst (4 entries per gate); recover them from the offsets
            offs = {it[2]: k for k, it in enumerate(tabs.items)}
            idx = [offs[int(o)] for o in seg.dst[::4].tolist()]
            qs = [cc.ops[i].qubits[0] for i in idx]
            assert len(set(qs)) == len(qs) and all(cc.ops[i].k == 1 for i in idx)
            kinds.add("rounds")
        else:
            idx = list(range(seg.first, seg.last))
            if isinstance(seg, autograd._DiagRun):
                assert all(cc.ops[i].kind[0] == "diag" for i in idx) and len(idx) >= autograd.diag_run_min
                kinds.add("D")
            elif isinstance(seg, autograd._OneQubitRun):
                qs = [cc.ops[i].qubits[0] for i in idx]
                assert len(set(qs)) == len(qs) and all(cc.ops[i].k == 1 for i in idx)
                kinds.add("S")
            else:
                assert isinstance(seg, autograd._ConstRun) and all(const[i] for i in idx)
                kinds.add("C")
        seen += idx
    assert sorted(seen) == list(range(len(structure)))
    assert kinds == {"G", "rounds", "D", "S", "C"}
    # program order is respected between segments that do not commute trivially: contiguous segments are ascending
    firsts = [s[1] if isinstance(s, tuple) else s.first for s in tabs.segments if not (not isinstance(s, tuple) and s.first == -1)]
    assert firsts == sorted(firsts)
    assert len(tabs.fused) >= 1  # the first rx layer follows an rzz layer? no: H layer (one-qubit run) + rzz run
    for k, fu in tabs.fused.items():
        assert isinstance(tabs.segments[k], autograd._OneQubitRun) and isinstance(tabs.segments[k + 1], autograd._DiagRun)


def test_backend_vmap_and_vvag_paths_on_plain_torch_functions():
    """backend.vmap / vvag semantics (pytorch_backend.py:816-878) independent of the engine: one torch.vmap
    evaluation when the function can be traced, the per-sample loop otherwise, identical results either way."""
    import torch

    from tensorcircuit_ng_b200 import backend

    x = torch.arange(6.0).reshape(3, 2)
    w = torch.tensor([1.0, 2.0])
    f = lambda a, b: (a * b).sum() ** 2  # noqa: E731
    old = backend.batched_mode
    try:
        backend.batched_mode = "strict"
        out = backend.vmap(f, vectorized_argnums=0)(x, w)
        assert backend.last_vmap_path == "batched" and out.tolist() == [4.0, 64.0, 196.0]
        v, (gx, gw) = backend.vvag(f, argnums=(0, 1), vectorized_argnums=0)(x, w)
        backend.batched_mode = "loop"
        v2, (gx2, gw2) = backend.vvag(f, argnums=(0, 1), vectorized_argnums=0)(x, w)
        assert backend.last_vmap_path.startswith("loop")
        assert torch.allclose(v, v2) and torch.allclose(gx, gx2) and torch.allclose(gw, gw2)
        assert tuple(gx.shape) == (3, 2) and tuple(gw.shape) == (2,)  # vectorised arg: stacked; shared arg: summed
        backend.batched_mode = "auto"
        h = lambda a: torch.tensor(float(a.sum()) * 2.0)  # noqa: E731  (data-dependent python: not traceable)
        out = backend.vmap(h)(x)
        assert backend.last_vmap_path.startswith("loop") and out.tolist() == [2.0, 10.0, 18.0]
        with pytest.raises(Exception):
            backend.batched_mode = "strict"
            backend.vmap(h)(x)
        backend.batched_mode = "auto"
        pair = backend.vmap(lambda a: (a.sum(), a.prod()))(x)  # tuple outputs
        assert pair[0].tolist() == [1.0, 5.0, 9.0] and pair[1].tolist() == [0.0, 6.0, 20.0]
        va, ga = backend.vvag(lambda a: (a.sum() ** 2, a.mean()), has_aux=True)(x)  # aux outputs take the loop
        assert backend.last_vmap_path.startswith("loop") and tuple(ga.shape) == (3, 2)
    finally:
        backend.batched_mode = old


def test_public_entry_points_fail_loudly_without_a_gpu(built):
    """No CPU fallback anywhere on the product path: without a CUDA device the statevector route, the torch
    interface, the Pauli-sum operators and the sampler raise EngineError instead of computing on the host."""
    import torch

    import tensorcircuit_ng_b200 as tc

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    c = tc.Circuit(2)
    c.h(0)
    c.cnot(0, 1)
    with pytest.raises(tc._lib.EngineError):
        c.wavefunction()
    with pytest.raises(tc._lib.EngineError):
        tc.interfaces.torch_interface(lambda p: p.sum())(torch.ones(2))
    h = tc.quantum.PauliStringSum([[3, 3]], [1.0])
    with pytest.raises(tc._lib.EngineError):
        h.expectation(torch.ones(4, dtype=torch.complex64))
    with pytest.raises(tc._lib.EngineError):
        tc.sampling.StateSampler(torch.ones(4, dtype=torch.complex64), 2)
    with pytest.raises(tc._lib.EngineError):
        c.sample(batch=2, allow_state=True)


def test_circuit_drop_frees_the_network_without_the_cycle_collector(built):
    """Node <-> Edge cycles are cut in `Circuit.__del__`: reference counting alone frees the ~4000 objects a
    630-gate circuit builds (the cyclic collector's full passes cost 150-200 ms with torch loaded)."""
    import gc
    import weakref

    import tensorcircuit_ng_b200 as tc

    gc.disable()
    try:
        c = tc.Circuit(6)
        for q in range(6):
            c.h(q)
        for q in range(5):
            c.rzz(q, q + 1, theta=torch.tensor(0.3))
            c.rx(q, theta=torch.tensor(0.7))
        refs = [weakref.ref(nd) for nd in c._nodes]  # (Edge has __slots__ without __weakref__; nodes tell the story)
        nodes, edges = c._copy()  # copies are independent of the circuit's own nodes
        del c
        assert all(r() is None for r in refs)
        assert len(nodes) == 6 + 6 + 10 and all(len(nd.edges) > 0 for nd in nodes)
    finally:
        gc.enable()


def test_deferred_gates_under_an_active_default_device_mode(built):
    """`torch.set_default_device` (the usual set-up next to the reference) installs a torch-function mode; the
    deferred-gate constructors bypass it (they only read attributes) and must build the same gates."""
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import gates

    th = torch.tensor(0.37, dtype=torch.float32)
    ref = gates.rx_gate(theta=th).tensor.clone()
    ref_zz = gates.rzz_gate(theta=th).tensor.clone()
    with torch.device("cpu"):  # DeviceContext mode active
        c = tc.Circuit(2)
        c.rx(0, theta=th)
        c.rzz(0, 1, theta=th)
        g_rx, g_zz = c._nodes[-2], c._nodes[-1]
        assert g_rx.pending() and g_zz.pending()
        assert torch.allclose(g_rx.tensor, ref) and torch.allclose(g_zz.tensor, ref_zz)
