"""GPU tier, round 2: boundary options, foreign nodes, gradient routes the adjoint method does not cover, and
full-width checks of the statevector path (n = 26 against the oracle, n = 30 against a closed form)."""
import numpy as np
import pytest
import torch

import tc_oracle
from helpers import brickwork, build, oracle_state, qaoa, random_layers

pytestmark = pytest.mark.gpu


def _tc():
    import tensorcircuit_ng_b200 as tc

    return tc


# ---- set_contractor options through the tensor-network engine ------------------------------------
@pytest.mark.parametrize("method,kw", [("greedy", {"preprocessing": True}), ("greedy", {}), ("eager", {}),
                                        ("auto", {"preprocessing": True}), ("tn", {})])
def test_named_contractors_match_oracle(cuda, method, kw):
    tc = _tc()
    n = 8
    ops = random_layers(n, 2, 4)
    ref = oracle_state(n, ops)
    with tc.runtime_contractor(method, **kw):
        psi = build(tc, n, ops).wavefunction()
        e = build(tc, n, ops).expectation_ps(z=[0, 3], x=[5])
    assert np.abs(psi.cpu().numpy().reshape(-1) - ref).max() <= 1e-5
    from tc_oracle import cons as ocons

    with ocons.runtime_contractor("greedy", preprocessing=True):
        eo = build(tc_oracle, n, ops).expectation_ps(z=[0, 3], x=[5])
    assert abs(complex(e.cpu()) - complex(eo)) <= 1e-5


def test_merge_single_gates_equals_oracle_and_custom_path(cuda):
    """`preprocessing=True`: same merged network as the oracle (node count, symbols), so a literal path computed
    for the reference's network contracts to the same state here (cons.py:1034-1040)."""
    tc = _tc()
    from tc_oracle import cons as ocons, paths
    from tensorcircuit_ng_b200 import cons

    n = 6
    ops = brickwork(n, 3, seed=2)
    pn = build(tc, n, ops)._copy()[0]
    on = build(tc_oracle, n, ops)._copy()[0]
    pm = cons._merge_single_gates(pn)
    om, _ = ocons._merge_single_gates(on)
    assert len(pm) == len(om)
    (pi, po, ps), _ = cons.get_tn_info(pm)
    (oi, oo, os_), _ = ocons.get_tn_info(om)
    assert (pi, po, ps) == (oi, oo, os_)
    path = paths.greedy(oi, oo, os_)  # the ORACLE's path on the ORACLE's merged network ...
    ref = oracle_state(n, ops)
    with tc.runtime_contractor("custom", optimizer=[tuple(p) for p in path], preprocessing=True):
        psi = build(tc, n, ops).wavefunction()  # ... applied by the product to its own merged network
    assert np.abs(psi.cpu().numpy().reshape(-1) - ref).max() <= 1e-5


def test_strip_exponent_and_contraction_info(cuda, capsys):
    tc = _tc()
    from tensorcircuit_ng_b200 import cons

    n = 6
    ops = brickwork(n, 2, seed=3)
    nodes = build(tc, n, ops).amplitude_before("010011")
    cf = cons.set_contractor("greedy", strip_exponent=True, contraction_info=True, set_global=False)
    node, exponent = cf(nodes)
    out = capsys.readouterr().out
    assert "log10[FLOPs]" in out
    ref = oracle_state(n, ops)[int("010011", 2)]
    val = complex(node.tensor.cpu()) * 10.0 ** float(exponent)
    assert abs(val - ref) <= 1e-6
    assert abs(abs(complex(node.tensor.cpu())) - 1.0) <= 1e-5  # mantissa normalised


# ---- foreign nodes ---------------------------------------------------------------------------------
def test_foreign_nodes_take_the_fast_paths(cuda):
    """Plain nodes carrying only reference-style names (no `_b200_kind`): same state, same number of fused passes
    as the native route; one probe readback for the topology."""
    tc = _tc()
    from test_boundary_host import _foreignize
    from tensorcircuit_ng_b200 import cons, svengine

    n = 14
    ops, _ = qaoa(n, 2, seed=1)
    ops = ops + [("cnot", [0, 5], {}), ("crx", [3, 9], {"theta": 0.7}), ("any", [2], {"unitary": np.diag([1.0, 1j])}),
                 ("rzz", [1, 12], {"theta": 0.3}), ("toffoli", [4, 6, 8], {})]  # fmt: skip
    ref = oracle_state(n, ops)

    def run(foreign):
        c = build(tc, n, ops)
        nodes, edges = c._copy()
        if foreign:
            _foreignize(nodes)
        gates = svengine.extract_gate_stream(nodes, edges)[2]
        structure = [(g[1], k, int(np.prod(g[0].shape))) for g, k in zip(gates, svengine.gate_kinds(gates))]
        cc = svengine.compile_circuit(n, structure, torch.device("cuda:0"), absorb_prefix=True)
        out = cons.b200_contractor(nodes, edges)
        return out.tensor.cpu().numpy().reshape(-1), cc.plan.n_launches

    svengine._probe_cache.clear()
    r0 = svengine.probe_readbacks
    psi_n, launches_n = run(False)
    psi_f, launches_f = run(True)
    assert svengine.probe_readbacks == r0 + 1
    assert np.abs(psi_n - ref).max() <= 1e-5 and np.abs(psi_f - ref).max() <= 1e-5
    assert launches_f == launches_n


# ---- gradients outside the adjoint method ---------------------------------------------------------
def _dense_reference(n, apply_ops, dtype=torch.complex128):
    """Plain torch statevector in complex128 (independent of the engine): apply_ops(state_fn) -> value."""

    def apply(psi, m, qs):
        k = len(qs)
        psi = psi.reshape([2] * n)
        m = m.reshape([2] * (2 * k)).to(dtype)
        psi = torch.tensordot(m, psi, dims=(list(range(k, 2 * k)), list(qs)))
        return torch.movedim(psi, list(range(k)), list(qs)).reshape(-1)

    psi = torch.zeros(2**n, dtype=dtype)
    psi[0] = 1
    return apply_ops(psi, apply)


def test_gradient_through_non_unitary_gate_matches_dense_reference(cuda):
    """the adjoint backward un-computes with U^dagger, which is wrong for a non-unitary `any` gate; such
    circuits are routed to the tensor-network path (ordinary autograd).  Value and gradient vs a complex128 torch
    statevector."""
    tc = _tc()
    n = 4
    rng = np.random.default_rng(3)
    m_np = (np.eye(2) + 0.3 * (rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))).astype(np.complex64)
    th0 = np.array([0.4, 1.1, -0.7], dtype=np.float32)

    def f_engine(th, m):
        c = tc.Circuit(n)
        for q in range(n):
            c.h(q)
        c.rx(0, theta=th[0])
        c.any(1, unitary=m)
        c.rzz(1, 2, theta=th[1])
        c.ry(3, theta=th[2])
        c.cnot(0, 3)
        return c.expectation_ps(z=[1, 2]).real + 0.5 * c.expectation_ps(x=[3]).real

    th = torch.tensor(th0, requires_grad=True, device="cuda")
    m = torch.tensor(m_np, requires_grad=True, device="cuda")
    val = f_engine(th, m)
    val.backward()

    def f_dense(th_, m_):
        from tensorcircuit_ng_b200 import gates as G

        def ops(psi, apply):
            H = torch.tensor(G._h_matrix, dtype=torch.complex128)
            X = torch.tensor(G._x_matrix, dtype=torch.complex128)
            Y = torch.tensor(G._y_matrix, dtype=torch.complex128)
            Z = torch.tensor(G._z_matrix, dtype=torch.complex128)
            I2 = torch.eye(2, dtype=torch.complex128)
            for q in range(n):
                psi = apply(psi, H, [q])
            psi = apply(psi, torch.cos(th_[0] / 2) * I2 - 1j * torch.sin(th_[0] / 2) * X, [0])
            psi = apply(psi, m_, [1])
            ZZ = torch.kron(Z, Z)
            psi = apply(psi, torch.cos(th_[1] / 2) * torch.eye(4, dtype=torch.complex128) - 1j * torch.sin(th_[1] / 2) * ZZ, [1, 2])
            psi = apply(psi, torch.cos(th_[2] / 2) * I2 - 1j * torch.sin(th_[2] / 2) * Y, [3])
            psi = apply(psi, torch.tensor(G._cnot_matrix, dtype=torch.complex128), [0, 3])
            zz = torch.vdot(psi, apply(apply(psi, Z, [1]), Z, [2])).real
            xx = torch.vdot(psi, apply(psi, X, [3])).real
            return zz + 0.5 * xx

        return _dense_reference(n, ops)

    th_r = torch.tensor(th0.astype(np.float64), requires_grad=True)
    m_r = torch.tensor(m_np.astype(np.complex128), requires_grad=True)
    ref = f_dense(th_r, m_r)
    ref.backward()
    assert abs(float(val) - float(ref)) <= 1e-5
    assert np.abs(th.grad.cpu().numpy() - th_r.grad.cpu().numpy()).max() <= 1e-4 * max(1.0, float(th_r.grad.abs().max()))
    assert np.abs(m.grad.cpu().numpy() - m_r.grad.cpu().numpy()).max() <= 1e-4 * max(1.0, float(m_r.grad.abs().max()))


def test_gradient_with_respect_to_operator_and_pauli_weights(cuda):
    """`expectation((w * Z, [0]))` with a trainable operator and `PauliStringSum` with trainable weights
    must carry their gradients (reference: ordinary autograd through the operator tensors)."""
    tc = _tc()
    n = 5
    ops = brickwork(n, 2, seed=6)
    ref = oracle_state(n, ops).astype(np.complex128)
    idx = np.arange(2**n)
    z0 = 1 - 2 * ((idx >> (n - 1)) & 1)
    want_z0 = float(np.sum(np.abs(ref) ** 2 * z0))
    w = torch.tensor(0.7, requires_grad=True, device="cuda")
    c = build(tc, n, ops)
    val = c.expectation((w * tc.gates.z().tensor, [0])).real
    val.backward()
    assert abs(float(val) - 0.7 * want_z0) <= 1e-5
    assert abs(float(w.grad) - want_z0) <= 1e-5
    # Pauli sum with live weights: dE/dw_t = <P_t>
    structures = [[3, 0, 0, 0, 0], [0, 1, 0, 0, 0], [3, 3, 0, 0, 0]]
    ws = torch.tensor([0.5, -1.2, 0.3], requires_grad=True, device="cuda")
    h = tc.quantum.PauliStringSum(structures, ws)
    psi = build(tc, n, ops).wavefunction()
    e = h.expectation(psi).real
    e.backward()
    z1 = 1 - 2 * ((idx >> (n - 2)) & 1)
    x1 = float(np.real(np.vdot(ref, ref.reshape([2] * n)[:, ::-1].reshape(-1))))
    terms = [want_z0, x1, float(np.sum(np.abs(ref) ** 2 * z0 * z1))]
    assert np.allclose(ws.grad.cpu().numpy(), terms, atol=1e-5)
    assert abs(float(e) - float(np.dot([0.5, -1.2, 0.3], terms))) <= 1e-5


# ---- full-width checks -------------------------------------------------------------------------------
def test_qaoa_n26_p2_matches_oracle(cuda):
    """nothing pinned the wide-state code paths to an independent value.  26 qubits, p = 2, against the
    numpy oracle (plain contractor order): amplitudes on a strided sample + every <Z_i Z_j> cost term."""
    tc = _tc()
    n, p = 26, 2
    ops, edges = qaoa(n, p, seed=0)
    ref = oracle_state(n, ops, contractor="plain").astype(np.complex64)
    c = build(tc, n, ops)
    psi = c.wavefunction().reshape(-1)
    sel = torch.arange(0, 2**n, 4099, device=psi.device)
    got = psi[sel].cpu().numpy()
    assert np.abs(got - ref[sel.cpu().numpy()]).max() <= 1e-5
    assert abs(float(torch.linalg.vector_norm(psi)) - 1.0) <= 1e-5
    prob = np.abs(ref.astype(np.complex128)) ** 2
    idx = np.arange(2**n, dtype=np.int64)
    for a, b in edges[::4]:
        sgn = (1 - 2 * ((idx >> (n - 1 - a)) & 1)) * (1 - 2 * ((idx >> (n - 1 - b)) & 1))
        want = float(np.sum(prob * sgn))
        assert abs(float(c.expectation_ps(z=[a, b]).real) - want) <= 1e-5


def test_qaoa_n30_p1_matches_closed_form(cuda):
    """30 qubits (byte offsets beyond 2^32, 2^18 tiles, the low_bits = 2 second plan): p = 1 MaxCut QAOA has a
    closed form for <Z_u Z_v> on every edge in terms of degrees and common neighbours (Wang, Hadfield, Jiang,
    Rieffel 2018, eq. 14) — an oracle that needs no 2^30 numpy state."""
    import networkx as nx

    tc = _tc()
    n = 30
    g = nx.random_regular_graph(3, n, seed=0)
    gamma, beta = 0.6, 0.35
    c = tc.Circuit(n)
    for q in range(n):
        c.h(q)
    for a, b in g.edges:
        c.exp1(int(a), int(b), unitary=tc.gates._zz_matrix, theta=gamma)  # exp(-i gamma Z Z)
    for q in range(n):
        c.rx(q, theta=2 * beta)  # exp(-i beta X)
    # the paper's unitaries are exp(-i gamma' C) exp(-i beta B) with C = sum (1 - ZZ) / 2: exp(-i gamma ZZ) here is
    # gamma' = 2 gamma up to a phase and a sign of the angle, which flips the sign of <ZZ> (checked against the
    # numpy oracle at n = 10 when this test was written)
    gp = 2 * gamma
    for u, v in list(g.edges)[::3]:
        d, e = g.degree[u] - 1, g.degree[v] - 1
        f = len(set(g[u]) & set(g[v]))
        zz = (-0.5 * np.sin(4 * beta) * np.sin(gp) * (np.cos(gp) ** d + np.cos(gp) ** e)
              - 0.5 * np.sin(2 * beta) ** 2 * np.cos(gp) ** (d + e - 2 * f) * (1 - np.cos(2 * gp) ** f))  # fmt: skip
        got = float(c.expectation_ps(z=[int(u), int(v)]).real)
        assert abs(got + zz) <= 2e-5, (u, v, got, -zz)


def test_json_round_trip_state_matches_oracle(cuda, tmp_path):
    """(f)3: `to_json -> from_json(_file)` and `from_qir` give the oracle's state of the original gate list."""
    tc = _tc()
    from test_boundary_host import _format_circuit

    n, ops, c = _format_circuit(tc)
    oc = build(tc_oracle, n, ops)
    oc.diagonal(0, 1, diag=np.exp(1j * np.arange(4)))
    from tc_oracle import cons as ocons

    with ocons.runtime_contractor("plain"):
        ref = np.asarray(oc.wavefunction()).reshape(-1)
    f = tmp_path / "c.json"
    c.to_json(file=str(f))
    for c2 in (tc.Circuit.from_json(c.to_json(simplified=True)), tc.Circuit.from_json_file(str(f)),
               tc.Circuit.from_qir(c.to_qir())):  # fmt: skip
        assert np.abs(c2.wavefunction().cpu().numpy().reshape(-1) - ref).max() <= 1e-5


def test_copynode_network_through_tn_route_matches_oracle(cuda):
    """R12: a hyperedge (CopyNode) network contracted by the tensor-network route (`use_primitives` path of
    cons.py:898-908) — state and an expectation sandwich built with `reuse=False` — against the oracle."""
    tc = _tc()

    def circ(mod):
        c = mod.Circuit(5)
        for q in range(5):
            c.h(q)
        c.diagonal(0, 1, diag=np.exp(1j * np.arange(4)).astype(np.complex64))
        c.rx(1, theta=0.3)
        c.diagonal(2, diag=np.array([1.0, 1j], dtype=np.complex64))
        c.cnot(2, 3)
        c.diagonal(3, 4, 0, diag=np.exp(0.5j * np.arange(8)).astype(np.complex64))
        c.ry(4, theta=-0.8)
        return c

    from tc_oracle import cons as ocons

    with ocons.runtime_contractor("plain"):
        ref = np.asarray(circ(tc_oracle).wavefunction()).reshape(-1)
        eref = complex(circ(tc_oracle).expectation_ps(z=[1], x=[4]))
    for method, kw in (("tn", {}), ("greedy", {}), ("custom", {"optimizer": None})):
        with tc.runtime_contractor(method, **kw):
            psi = circ(tc).wavefunction()
            e = circ(tc).expectation((tc.gates.z(), [1]), (tc.gates.x(), [4]), reuse=False)
        assert np.abs(psi.cpu().numpy().reshape(-1) - ref).max() <= 1e-5, method
        assert abs(complex(e.cpu()) - eref) <= 1e-5, method
