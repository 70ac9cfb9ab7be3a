"""GPU tier: the CUDA path (through the C ABI) against the numpy oracle on identical circuits.
Tolerances are the north-star's: max|dpsi| <= 1e-5 (complex64), expectations 1e-5, gradients 1e-4 rel."""
import numpy as np
import pytest
import torch

import tc_oracle
from helpers import brickwork, build, oracle_circuit, oracle_state, qaoa, random_layers

pytestmark = pytest.mark.gpu
ATOL_PSI = 1e-5


def _tc():
    import tensorcircuit_ng_b200 as tc

    return tc


def _np(t):
    return t.detach().cpu().numpy()


# ---- reference golden values through the CUDA path -------------------------------------------------
def test_reference_kats_on_gpu(cuda):
    tc = _tc()
    g = lambda s: tc.gates.Gate(torch.arange(s, dtype=torch.float32).reshape([2] * int(np.log2(s))).to(torch.complex64))
    qc = tc.Circuit(2)
    qc.unitary(0, 1, unitary=g(16))
    assert _np(qc.wavefunction())[2].real == 8
    qc = tc.Circuit(2)
    qc.unitary(1, 0, unitary=g(16))
    assert _np(qc.wavefunction())[2].real == 4
    qc = tc.Circuit(2)
    qc.unitary(0, unitary=g(4))
    assert _np(qc.wavefunction())[2].real == 2
    c = tc.Circuit(2)
    c.x(1)
    c.crx(1, 0, theta=0.3)
    np.testing.assert_allclose(_np(c.expectation([tc.gates._z_matrix, 0])), 0.95533645, atol=1e-5)
    c = tc.Circuit(1)
    c.X(0)
    c.SD(0)
    np.testing.assert_allclose(_np(c.state()), np.array([0.0, -1.0j]), atol=1e-6)
    c = tc.Circuit(2)
    c.X(0)
    np.testing.assert_allclose(_np(c.expectation_ps(z=[0, 1])), -1, atol=1e-5)
    c = tc.Circuit(2)
    c.H(0)
    np.testing.assert_allclose(_np(c.expectation_ps(z=[1], x=[0])), 1, atol=1e-5)
    np.testing.assert_allclose(_np(c.expectation_ps(ps=[1, 3])), 1, atol=1e-5)
    c = tc.Circuit(1, inputs=1 / np.sqrt(2) * np.array([-1, 1.0j]))
    np.testing.assert_allclose(_np(c.expectation_ps(y=[0])), -1, atol=1e-5)
    c = tc.Circuit(3)
    for _ in range(2):
        c.H(0)
        c.rx(1, theta=0.7)
        c.exp1(0, 1, unitary=tc.gates._zz_matrix, theta=-0.2)
    np.testing.assert_allclose(_np(c.expectation((tc.gates.z(), [1]))), 0.202728, atol=1e-5)
    c = tc.Circuit(2, inputs=np.eye(4))
    c.X(0)
    c.Y(1)
    np.testing.assert_allclose(_np(c.wavefunction()).reshape(4, 4), np.kron(tc.gates._x_matrix, tc.gates._y_matrix), atol=1e-4)
    c = tc.Circuit(2)
    c.iswap(0, 1, theta=-0.2)
    c.cphase(0, 1, theta=-0.3)
    ans = np.array([[1.0, 0, 0, 0], [0, 0.95105654, -0.309017j, 0], [0, -0.309017j, 0.95105654, 0],
                    [0, 0, 0, 0.9553365 - 0.29552022j]])  # fmt: skip
    np.testing.assert_allclose(_np(c.matrix()), ans, atol=1e-5)
    c = tc.Circuit(2)
    c.x(0)
    np.testing.assert_allclose(_np(c.amplitude("10")), 1.0, atol=1e-6)
    c.CNOT(0, 1)
    np.testing.assert_allclose(_np(c.amplitude("11")), 1.0, atol=1e-6)


# ---- statevector parity ------------------------------------------------------------------------------
@pytest.mark.parametrize("n,depth,seed", [(1, 2, 0), (2, 3, 1), (5, 3, 2), (9, 3, 3), (10, 3, 4), (12, 4, 5),
                                          (13, 3, 6), (14, 3, 7), (17, 2, 8), (20, 2, 9)])  # fmt: skip
def test_random_circuit_state(cuda, n, depth, seed):
    tc = _tc()
    ops = random_layers(n, depth, seed)
    psi = _np(build(tc, n, ops).wavefunction())
    ref = oracle_state(n, ops)
    assert np.abs(psi - ref).max() <= ATOL_PSI


@pytest.mark.parametrize("tile_bits,low_bits", [(10, 4), (11, 3), (12, 5), (13, 4), (13, 5)])
def test_tile_geometries(cuda, tile_bits, low_bits):
    tc = _tc()
    n = 16
    ops = random_layers(n, 2, 21)
    old = dict(tc.svengine.plan_options)
    tc.svengine.plan_options.update(tile_bits=tile_bits, low_bits=low_bits)
    try:
        psi = _np(build(tc, n, ops).wavefunction())
    finally:
        tc.svengine.plan_options.clear()
        tc.svengine.plan_options.update(old)
    assert np.abs(psi - oracle_state(n, ops)).max() <= ATOL_PSI


def test_config1_brickwork_20q(cuda):
    """BASELINE.json configs[0]: 20-qubit brickwork depth 10, wavefunction + <Z0Z1>."""
    tc = _tc()
    n = 20
    ops = brickwork(n, 10)
    c = build(tc, n, ops)
    psi = _np(c.wavefunction())
    e = complex(_np(c.expectation_ps(z=[0, 1])))
    from tc_oracle import cons

    with cons.runtime_contractor("greedy", preprocessing=True):  # the reference's default contractor
        co = oracle_circuit(n, ops)
        ref = co.wavefunction()
        eref = complex(co.expectation_ps(z=[0, 1]))
    assert np.abs(psi - ref).max() <= ATOL_PSI
    assert abs(e - eref) <= 1e-5


def test_qaoa_maxcut_state_and_cost(cuda):
    """BASELINE.json configs[2] at an oracle-checkable width (18 qubits, p = 4)."""
    tc = _tc()
    n, p = 18, 4
    ops, edges = qaoa(n, p)
    c = build(tc, n, ops)
    psi = _np(c.wavefunction())
    ref = oracle_state(n, ops)
    assert np.abs(psi - ref).max() <= ATOL_PSI
    cost = sum(0.5 * (1 - complex(_np(c.expectation_ps(z=[a, b]))).real) for a, b in edges)
    probs = np.abs(ref.astype(np.complex128)) ** 2
    idx = np.arange(2**n)
    cref = 0.0
    for a, b in edges:
        za = 1 - 2 * ((idx >> (n - 1 - a)) & 1)
        zb = 1 - 2 * ((idx >> (n - 1 - b)) & 1)
        cref += 0.5 * (1 - float(np.sum(probs * za * zb)))
    assert abs(cost - cref) <= 1e-5 * max(1.0, abs(cref))
    # batched Z-string kernel: all edges in one read of the state
    zz = _np(tc.expect.z_expectations(c.wavefunction(), n, [[a, b] for a, b in edges]))
    assert abs(float(np.sum(0.5 * (1 - zz))) - cref) <= 1e-5 * max(1.0, abs(cref))


def test_inputs_and_matrix(cuda):
    tc = _tc()
    n = 11
    rng = np.random.default_rng(3)
    v = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    v = (v / np.linalg.norm(v)).astype(np.complex64)
    ops = random_layers(n, 2, 33)
    psi = _np(build(tc, n, ops, inputs=v).wavefunction())
    ref = oracle_state(n, ops, inputs=v, contractor="plain")
    assert np.abs(psi - ref).max() <= ATOL_PSI
    ops4 = random_layers(4, 2, 5)
    m = _np(build(tc, 4, ops4).matrix())
    mref = oracle_circuit(4, ops4).matrix()
    assert np.abs(m - mref).max() <= ATOL_PSI


def test_full_width_properties(cuda):
    """Size-independent properties at a width the oracle cannot check quickly (26 qubits):
    norm preservation and U^dagger U = 1 (circuit followed by its inverse returns |0>)."""
    tc = _tc()
    n = 26
    ops = brickwork(n, 4, seed=5)
    c = build(tc, n, ops)
    psi = c.wavefunction()
    assert abs(float(torch.linalg.vector_norm(psi)) - 1.0) <= 1e-4
    inv = [(name, qs, {"theta": -kw["theta"]}) for name, qs, kw in reversed(ops)]
    back = build(tc, n, ops + inv).wavefunction()
    assert abs(abs(complex(back[0])) - 1.0) <= 1e-4
    assert float(back[1:].abs().max()) <= 1e-4


# ---- expectations -----------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 9, 14])
def test_expectation_ps_all_paulis(cuda, n):
    tc = _tc()
    ops = random_layers(n, 2, 40 + n)
    c = build(tc, n, ops)
    co = oracle_circuit(n, ops)
    from tc_oracle import cons

    with cons.runtime_contractor("plain"):
        for kw in [{"z": [0]}, {"x": [n - 1]}, {"y": [1]}, {"x": [0], "z": [n - 1]}, {"x": [0], "y": [1], "z": [2]},
                   {"z": [0, n - 1]}, {"y": [0, 2]}]:  # fmt: skip
            got = complex(_np(c.expectation_ps(**kw)))
            want = complex(co.expectation_ps(**kw))
            assert abs(got - want) <= 1e-5, kw


def test_general_operator_expectation(cuda):
    tc = _tc()
    n = 8
    ops = random_layers(n, 2, 77)
    c = build(tc, n, ops)
    co = oracle_circuit(n, ops)
    rng = np.random.default_rng(1)
    m1 = (rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))).astype(np.complex64)
    m2 = (rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))).astype(np.complex64)
    got = complex(_np(c.expectation([m1, [2]], [m2, [5, 0]])))
    want = complex(co.expectation([m1, [2]], [m2, [5, 0]]))
    assert abs(got - want) <= 1e-5 * max(1, abs(want))
    got = complex(_np(c.expectation([m1, [2]], reuse=False)))
    want = complex(co.expectation([m1, [2]], reuse=False))
    assert abs(got - want) <= 1e-5 * max(1, abs(want))


def test_lightcone_expectation(cuda):
    tc = _tc()

    def construct_c(mod, pbc=True):
        n = 4
        ns = n if pbc else n - 1
        c = mod.Circuit(n)
        for j in range(2):
            for i in range(n):
                c.rx(i, theta=0.2, name="rx" + str(j) + "-" + str(i))
            for i in range(ns):
                c.cnot(i, (i + 1) % n, name="cnot" + str(j) + "-" + str(i))
        return c

    for b in [True, False]:
        c = construct_c(tc, b)
        m1 = complex(_np(c.expectation_ps(z=[0], enable_lightcone=True)))
        m2 = complex(_np(c.expectation_ps(z=[0])))
        want = complex(construct_c(tc_oracle, b).expectation_ps(z=[0]))
        assert abs(m1 - want) <= 1e-5 and abs(m2 - want) <= 1e-5
        nodes = c.expectation_before([tc.gates.z(), 0], reuse=False)
        l1 = len(nodes)
        l2 = len(tc.simplify._full_light_cone_cancel(nodes))
        assert (l1, l2) == ((41, 41) if b else (37, 25))


# ---- tensor-network route ------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["tn", "plain"])
def test_tn_route_state_matches(cuda, method):
    tc = _tc()
    n = 8
    ops = random_layers(n, 2, 55)
    with tc.runtime_contractor(method):
        psi = _np(build(tc, n, ops).wavefunction())
    assert np.abs(psi - oracle_state(n, ops)).max() <= ATOL_PSI


def test_amplitude_network(cuda):
    tc = _tc()
    n = 10
    ops = random_layers(n, 3, 66)
    ref = oracle_state(n, ops)
    c = build(tc, n, ops)
    for bits in ["0" * n, "1" * n, "0110100101"]:
        got = complex(_np(c.amplitude(bits)))
        assert abs(got - ref[int(bits, 2)]) <= ATOL_PSI
    got = complex(_np(c.amplitude(torch.tensor([0, 1, 1, 0, 1, 0, 0, 1, 0, 1]))))
    assert abs(got - ref[int("0110100101", 2)]) <= ATOL_PSI


def test_custom_optimizer_plug(cuda):
    """Level-1 plug: caller-supplied optimizer / literal path, our executor."""
    tc = _tc()
    from tc_oracle import paths

    n = 6
    ops = random_layers(n, 2, 88)
    ref = oracle_state(n, ops)
    seen = {}

    def opt(inputs, output, size_dict, memory_limit=None):
        seen["called"] = True
        return paths.greedy(inputs, output, size_dict)

    with tc.runtime_contractor("custom", optimizer=opt):
        psi = _np(build(tc, n, ops).wavefunction())
    assert seen.get("called") and np.abs(psi - ref).max() <= ATOL_PSI


def test_partial_contraction_keeps_original_edges(cuda):
    """tests/test_hyperedge.py:498-527 of the reference: the returned node carries the ORIGINAL Edge objects."""
    tc = _tc()
    a = tc.tn.Node(torch.randn(2, 2, 2, dtype=torch.complex64))
    b = tc.tn.Node(torch.randn(2, 2, dtype=torch.complex64))
    a[2] ^ b[0]
    e0, e1, e2 = a[0], a[1], b[1]
    at, bt = a.tensor.clone(), b.tensor.clone()
    with tc.runtime_contractor("tn"):
        r = tc.cons.contractor([a, b], output_edge_order=[e2, e0, e1])
    assert r.edges[0] is e2 and r.edges[1] is e0 and r.edges[2] is e1
    assert e2.node1 is r and e0.node1 is r
    want = torch.einsum("abk,kc->cab", at, bt)
    assert torch.allclose(r.tensor, want, atol=1e-5)


def test_diagonal_hyperedge_gate(cuda):
    """tests/test_hyperedge.py:530-559 of the reference: c.diagonal == dense any(diagflat(d))."""
    tc = _tc()
    for n in (3, 11):
        d = np.exp(1j * np.arange(4) * 0.3).astype(np.complex64)
        c1 = tc.Circuit(n)
        c2 = tc.Circuit(n)
        for c in (c1, c2):
            for q in range(n):
                c.h(q)
        c1.diagonal(0, 2, diag=d)
        c2.any(0, 2, unitary=np.diagflat(d))
        for c in (c1, c2):
            c.rx(1, theta=0.4)
        assert (c1.state() - c2.state()).abs().max() <= ATOL_PSI
        e1 = complex(_np(c1.expectation_ps(z=[0], y=[1])))
        e2 = complex(_np(c2.expectation_ps(z=[0], y=[1])))
        assert abs(e1 - e2) <= 1e-5


# ---- gradients -------------------------------------------------------------------------------------------
def _example_block(mod, n, param, nlayers):
    c = mod.Circuit(n)
    zz = tc_oracle.gates._zz_matrix
    for i in range(n):
        c.H(i)
    for j in range(nlayers):
        for i in range(n - 1):
            c.exp1(i, i + 1, unitary=zz, theta=param[2 * j, i])
        for i in range(n):
            c.rx(i, theta=param[2 * j + 1, i])
    return c


def test_gradient_kat(cuda):
    """tests/test_interfaces.py:28-58 of the reference: d(<X1>^2)/dp[0,1] = -2.146e-3."""
    tc = _tc()
    n = 4
    param = torch.ones([4, n], requires_grad=True)
    c = _example_block(tc, n, param, 2)
    loss = c.expectation([tc.gates.x(), [1]]).real ** 2
    loss.backward()
    assert param.grad.shape == (4, n)
    np.testing.assert_allclose(float(param.grad[0, 1]), -2.146e-3, atol=1e-5)


@pytest.mark.parametrize("n", [4, 11])
def test_gradients_match_parameter_shift_of_oracle(cuda, n):
    """north-star: gradients within 1e-4 relative.  The reference value of every derivative is the exact
    parameter-shift rule evaluated on the oracle (no finite-difference step error)."""
    tc = _tc()
    rng = np.random.default_rng(n)
    p0 = rng.uniform(0, 1, size=(4, n))

    def f_oracle(p):
        c = _example_block(tc_oracle, n, p, 2)
        from tc_oracle import cons

        with cons.runtime_contractor("plain"):
            return float(np.real(c.expectation_ps(z=[0, 1])) + 0.5 * np.real(c.expectation_ps(x=[n - 1])))

    param = torch.tensor(p0, dtype=torch.float32, requires_grad=True)
    c = _example_block(tc, n, param, 2)
    val = c.expectation_ps(z=[0, 1]).real + 0.5 * c.expectation_ps(x=[n - 1]).real
    val.backward()
    g = param.grad.cpu().numpy()
    assert abs(float(val) - f_oracle(p0)) <= 1e-5
    from helpers import param_shift

    checks = [(0, 0), (1, n - 1), (2, 1), (3, 0), (0, n - 2), (1, 0), (2, 0), (3, n - 1)]
    ps = {idx: param_shift(f_oracle, p0, idx, "full" if idx[0] % 2 == 0 else "half") for idx in checks}
    scale = max(abs(v) for v in ps.values())
    for idx in checks:  # rows 0, 2: exp1(ZZ, theta) = exp(-i theta ZZ); rows 1, 3: rx(theta)
        assert abs(g[idx] - ps[idx]) <= 1e-4 * max(abs(ps[idx]), scale), (idx, g[idx], ps[idx])


def test_tn_route_gradient(cuda):
    tc = _tc()
    n = 4
    param = torch.ones([4, n], requires_grad=True)
    with tc.runtime_contractor("tn"):
        c = _example_block(tc, n, param, 2)
        loss = c.expectation([tc.gates.x(), [1]], reuse=False).real ** 2
    loss.backward()
    np.testing.assert_allclose(float(param.grad[0, 1]), -2.146e-3, atol=1e-5)


def test_vvag_tfim_vqe_batch(cuda):
    """configs[1] at test size: TFIM hardware-efficient ansatz (examples/benchmark_jax_vs_torch_vqe.py:160-200),
    `vvag` over a batch of parameter sets; values vs the oracle, gradients vs central differences of the
    oracle (atol 1e-4 relative: tests/test_backends.py:916-941 style)."""
    import tensorcircuit_ng_b200 as tc

    n, depth, batch = 8, 2, 3
    rng = np.random.default_rng(5)
    params = rng.normal(0, 0.4, size=(batch, depth, 2, n)).astype(np.float32)

    def energy(mod, p, to_float):
        c = mod.Circuit(n)
        for q in range(n):
            c.h(q)
        for l in range(depth):
            for q in range(n - 1):
                c.rzz(q, q + 1, theta=p[l, 0, q])
            for q in range(n):
                c.rx(q, theta=p[l, 1, q])
        e = 0.0
        for q in range(n - 1):
            e = e - to_float(c.expectation_ps(z=[q, q + 1]))
        for q in range(n):
            e = e - to_float(c.expectation_ps(x=[q]))
        return e

    f = lambda p: energy(tc, p, lambda v: v.real)  # noqa: E731
    vals, grads = tc.backend.vvag(f, argnums=0, vectorized_argnums=0)(torch.from_numpy(params).cuda())
    assert vals.shape == (batch,) and grads.shape == params.shape
    ref = lambda p: float(energy(tc_oracle, p, lambda v: np.real(v)))  # noqa: E731
    for b in range(batch):
        assert abs(float(vals[b]) - ref(params[b])) <= 1e-4
    from helpers import param_shift

    checks = [(0, 0, 0, 0), (1, 1, 1, 3), (2, 0, 1, 7), (2, 1, 0, 5), (0, 1, 0, 6), (1, 0, 1, 0)]
    ps = {c: param_shift(ref, params[c[0]], c[1:], "half") for c in checks}  # rzz(theta), rx(theta): exp(-i theta/2 P)
    scale = max(abs(v) for v in ps.values())
    for c in checks:
        assert abs(float(grads[c]) - ps[c]) <= 1e-4 * max(abs(ps[c]), scale), (c, float(grads[c]), ps[c])


def test_vvag_batched_path_equals_loop(cuda):
    """backend.vvag / vmap: one evaluation under torch.vmap with the kernels launched at batch = B gives the
    values and per-sample gradients of the per-sample loop; functions outside the batched path fall back."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import backend

    n, depth, B = 12, 2, 5
    ls, ws = [], []
    for q in range(n - 1):
        t = [0] * n
        t[q] = t[q + 1] = 3
        ls.append(t)
        ws.append(-1.0)
    for q in range(n):
        t = [0] * n
        t[q] = 1
        ls.append(t)
        ws.append(-0.7)
    ham = tc.quantum.PauliStringSum2COO(ls, ws)

    def energy(p, shift):
        c = tc.Circuit(n)
        for q in range(n):
            c.h(q)
        for l in range(depth):
            for q in range(n - 1):
                c.rzz(q, q + 1, theta=p[l, 0, q] * 2.0)
            c.cnot(0, 1)
            for q in range(n):
                c.rx(q, theta=p[l, 1, q] + shift)
        return tc.templates.measurements.operator_expectation(c, ham)

    torch.manual_seed(1)
    p = 0.4 * torch.randn(B, depth, 2, n)
    shift = torch.tensor(0.3)
    old = backend.batched_mode
    try:
        backend.batched_mode = "strict"
        v1, g1 = backend.vvag(energy, argnums=0, vectorized_argnums=0)(p, shift)
        assert backend.last_vmap_path == "batched"
        v1s, (g1p, g1s) = backend.vvag(energy, argnums=(0, 1), vectorized_argnums=0)(p, shift)
        vm = backend.vmap(energy, vectorized_argnums=0)(p, shift)
        backend.batched_mode = "loop"
        v2, g2 = backend.vvag(energy, argnums=0, vectorized_argnums=0)(p, shift)
        v2s, (g2p, g2s) = backend.vvag(energy, argnums=(0, 1), vectorized_argnums=0)(p, shift)
        assert backend.last_vmap_path.startswith("loop")
        backend.batched_mode = "auto"

        def uses_item(x):  # data-dependent python: cannot be traced by vmap -> falls back to the loop
            c = tc.Circuit(3)
            c.rx(0, theta=float(x[0]))
            return c.expectation_ps(z=[0]).real

        out = backend.vmap(uses_item)(torch.tensor([[0.1], [0.2]]))
        assert backend.last_vmap_path.startswith("loop") and tuple(out.shape) == (2,)
    finally:
        backend.batched_mode = old
    assert tuple(v1.shape) == (B,) and tuple(g1.shape) == tuple(p.shape)
    assert float((v1 - v2).abs().max()) < 2e-5 and float((vm - v2).abs().max()) < 2e-5
    assert float((g1 - g2).abs().max()) < 5e-5
    assert float((g1p - g2p).abs().max()) < 5e-5 and abs(float(g1s) - float(g2s)) < 2e-4  # shared arg: summed
    assert float(g2.abs().max()) > 1e-2


def test_vmap_batches_expectation_ps_circuits(cuda):
    """The QML pattern (tests/test_torchnn.py:21-49 of the reference): per-qubit expectation_ps read-outs of a
    data-encoding circuit, batched over the data — one torch.vmap evaluation == the per-sample loop, values and
    weight gradients."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import backend

    n, nlayers, B = 6, 2, 4

    def qpred(x, weights):
        c = tc.Circuit(n)
        for i in range(n):
            c.rx(i, theta=x[i])
        for j in range(nlayers):
            for i in range(n - 1):
                c.cnot(i, i + 1)
            for i in range(n):
                c.rx(i, theta=weights[2 * j, i])
                c.ry(i, theta=weights[2 * j + 1, i])
        outs = [c.expectation_ps(x=[i]) for i in range(n)] + [c.expectation_ps(z=[0, 1]), c.expectation_ps(y=[2], z=[3])]
        return tc.backend.real(tc.backend.stack(outs))

    torch.manual_seed(3)
    xs = torch.rand(B, n)
    w = torch.randn(2 * nlayers, n)
    old = backend.batched_mode
    try:
        backend.batched_mode = "strict"
        y1 = backend.vmap(qpred, vectorized_argnums=0)(xs, w)
        loss = lambda xx, ww: qpred(xx, ww).sum()  # noqa: E731
        v1, g1 = backend.vvag(loss, argnums=1, vectorized_argnums=0)(xs, w)
        assert backend.last_vmap_path == "batched"
        backend.batched_mode = "loop"
        y2 = backend.vmap(qpred, vectorized_argnums=0)(xs, w)
        v2, g2 = backend.vvag(loss, argnums=1, vectorized_argnums=0)(xs, w)
    finally:
        backend.batched_mode = old
    assert tuple(y1.shape) == (B, n + 2)
    assert float((y1 - y2).abs().max()) < 1e-5
    assert float((v1 - v2).abs().max()) < 1e-5 and float((g1 - g2).abs().max()) < 5e-5
    assert float(g2.abs().max()) > 1e-2


def test_z_moment_tables_follow_circuit_couplings(cuda):
    """Repeated <Z_i>, <Z_i Z_j> queries on one state: the first is a direct reduction, the second builds the table
    of the circuit's own couplings, a pair outside it builds the full table — every answer equals the oracle's."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import cons

    n = 9
    rng = np.random.default_rng(11)
    th = rng.uniform(0, 2 * np.pi, size=(3, n))

    def build(mod):
        c = mod.Circuit(n)
        for q in range(n):
            c.h(q)
        for l in range(3):
            for q in range(n - 1):
                c.rzz(q, q + 1, theta=float(th[l, q]))
            for q in range(n):
                c.rx(q, theta=float(th[l, (q + 1) % n]))
        return c

    c, oc = build(tc), build(tc_oracle)
    queries = [[0, 1], [3, 4], [5], [7, 8], [2, 3], [0, 5], [1, 7], [4], [6, 8]]  # [0,5], [1,7], [6,8]: not coupled
    with torch.no_grad():
        for zq in queries:
            got = float(c.expectation_ps(z=zq).real)
            want = float(np.real(oc.expectation_ps(z=zq)))
            assert abs(got - want) < 2e-6, zq
    state = c._copy_state_tensor()[0][0].tensor
    src = c.state_tensor.tensor
    cache = getattr(src, "_b200_zcache", None)
    assert cache is not None and len(cache["tables"]) == 2
    assert len(cache["tables"][0][0]) == n + (n - 1) and len(cache["tables"][1][0]) == n + n * (n - 1) // 2
    assert cons.speculate_z_moments
