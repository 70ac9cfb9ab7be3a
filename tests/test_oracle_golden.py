"""The numpy oracle against the reference's own golden values.  CPU only."""
import json
import os

import numpy as np
import pytest

import tc_oracle as tc
from tc_oracle import circuit as ocirc
from tc_oracle import cons, gates, paths

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


@pytest.fixture(params=["greedy", "plain"])
def contractor(request):
    # every KAT must hold under the reference's default contractor and the literal statevector order
    with cons.runtime_contractor(request.param, **({"preprocessing": True} if request.param == "greedy" else {})):
        yield request.param


def test_wavefunction_arange(contractor):
    g = lambda s: gates.Gate(np.arange(s).reshape([2] * int(np.log2(s))).astype(np.complex64))
    qc = tc.Circuit(2)
    qc.unitary(0, 1, unitary=g(16))
    assert np.real(qc.wavefunction()[2]) == 8
    qc = tc.Circuit(2)
    qc.unitary(1, 0, unitary=g(16))
    assert np.real(qc.wavefunction()[2]) == 4
    qc = tc.Circuit(2)
    qc.unitary(0, unitary=g(4))
    assert np.real(qc.wavefunction()[2]) == 2


def test_basics(contractor):
    c = tc.Circuit(2)
    c.x(0)
    np.testing.assert_allclose(c.amplitude("10"), 1.0)
    c.CNOT(0, 1)
    np.testing.assert_allclose(c.amplitude("11"), 1.0)


def test_control_vgate(contractor):
    c = tc.Circuit(2)
    c.x(1)
    c.crx(1, 0, theta=0.3)
    k = KATS["crx_expectation"]
    np.testing.assert_allclose(c.expectation([gates._z_matrix, 0]), k["value"], atol=k["atol"])


def test_adjoint_gate(contractor):
    c = tc.Circuit(1)
    c.X(0)
    c.SD(0)
    np.testing.assert_allclose(c.state(), np.array([0.0, -1.0j]))


def test_expectations(contractor):
    c = tc.Circuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation((gates.z(), [0])), 0, atol=1e-7)
    c = tc.Circuit(2)
    c.X(0)
    np.testing.assert_allclose(c.expectation_ps(z=[0, 1]), -1, atol=1e-5)
    c = tc.Circuit(2)
    c.H(0)
    np.testing.assert_allclose(c.expectation_ps(z=[1], x=[0]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(ps=[1, 3]), 1, atol=1e-5)
    np.testing.assert_allclose(c.expectation_ps(z=[1, 2], ps=[1, 3]), 1, atol=1e-5)
    c = tc.Circuit(1, inputs=1 / np.sqrt(2) * np.array([-1, 1.0j]))
    np.testing.assert_allclose(c.expectation_ps(y=[0]), -1, atol=1e-5)


def test_unitary_and_iswap(contractor):
    c = tc.Circuit(2, inputs=np.eye(4))
    c.X(0)
    c.Y(1)
    np.testing.assert_allclose(c.wavefunction().reshape([4, 4]), np.kron(gates._x_matrix, gates._y_matrix), atol=1e-4)
    c = tc.Circuit(2, inputs=np.eye(2**2))
    c.iswap(0, 1)
    np.testing.assert_allclose(c.state().reshape([4, 4]), gates.iswap_gate().tensor.reshape([4, 4]), atol=1e-5)
    ans = np.array([[1.0, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1.0]])
    np.testing.assert_allclose(gates.iswap_gate().tensor, ans.reshape([2, 2, 2, 2]), atol=1e-5)
    np.testing.assert_allclose(gates.iswap_gate(theta=0).tensor, np.eye(4).reshape([2, 2, 2, 2]), atol=1e-5)


def test_toqir_value(contractor):
    # tests/test_circuit.py:856-876: the circuit applied twice => <Z1> = 0.202728
    c = tc.Circuit(3)
    for _ in range(2):
        c.H(0)
        c.rx(1, theta=0.7)
        c.exp1(0, 1, unitary=gates._zz_matrix, theta=-0.2)
    k = KATS["toqir_z1"]
    np.testing.assert_allclose(c.expectation((gates.z(), [1])), k["value"], atol=k["atol"])


def test_gate_kats(contractor):
    c = tc.Circuit(1)
    c.h(0)
    c.phase(0, theta=np.pi / 2)
    np.testing.assert_allclose(c.state()[1], 0.7071j, atol=1e-4)
    c = tc.Circuit(2)
    c.cu(0, 1, theta=np.pi / 2, phi=-np.pi / 4, lbd=np.pi / 4)
    m = c.matrix()
    np.testing.assert_allclose(m[2:, 2:], gates._wroot_matrix, atol=1e-5)
    np.testing.assert_allclose(m[:2, :2], np.eye(2), atol=1e-5)
    c = tc.Circuit(2)
    c.iswap(0, 1, theta=-0.2)
    c.cphase(0, 1, theta=-0.3)
    ans = np.array(
        [
            [1.0, 0, 0, 0],
            [0, 0.95105654, -0.309017j, 0],
            [0, -0.309017j, 0.95105654, 0],
            [0, 0, 0, 0.9553365 - 0.29552022j],
        ]
    )
    np.testing.assert_allclose(c.matrix(), ans, atol=1e-5)
    c = tc.Circuit(2)
    c.exp(0, 1, unitary=np.diag([1.0, -1, -1, 1]), theta=np.pi / 2)
    np.testing.assert_allclose(c.wavefunction()[0], -1j, atol=1e-6)
    c = tc.Circuit(2)
    c.any(0, unitary=np.eye(2))
    np.testing.assert_allclose(c.expectation((gates.z(), [0])), 1.0)


def test_rxx_ryy_rzz(contractor):
    c1 = tc.Circuit(3)
    c1.rxx(0, 1, theta=1.0)
    c1.ryy(0, 2, theta=0.5)
    c1.rzz(0, 1, theta=-0.5)
    c2 = tc.Circuit(3)
    c2.exp1(0, 1, theta=1.0 / 2, unitary=gates._xx_matrix)
    c2.exp1(0, 2, theta=0.5 / 2, unitary=gates._yy_matrix)
    c2.exp1(0, 1, theta=-0.5 / 2, unitary=gates._zz_matrix)
    np.testing.assert_allclose(c1.state(), c2.state(), atol=1e-5)


def _example_block(c, param, nlayers):
    n = c._nqubits
    param = np.reshape(param, [2 * nlayers, n])
    for i in range(n):
        c.H(i)
    for j in range(nlayers):
        for i in range(n - 1):
            c.exp1(i, i + 1, unitary=gates._zz_matrix, theta=param[2 * j, i])
        for i in range(n):
            c.rx(i, theta=param[2 * j + 1, i])
    return c


def test_gradient_kat():
    # tests/test_interfaces.py:28-58: d(<X1>^2)/dp[0,1] = -2.146e-3 at p = ones([4, 4]) (central difference)
    n = 4

    def f(p):
        c = _example_block(tc.Circuit(n), p, 2)
        return float(np.real(c.expectation([gates.x(), [1]]))) ** 2

    p = np.ones([4, n], dtype=np.float64)
    eps = 1e-3
    pp, pm = p.copy(), p.copy()
    pp[0, 1] += eps
    pm[0, 1] -= eps
    g = (f(pp) - f(pm)) / (2 * eps)
    k = KATS["torch_interface_grad"]
    np.testing.assert_allclose(g, k["value"], atol=5e-5)


def test_greedy_published_cost():
    """opt_einsum greedy restatement reproduces the published plan cost of the reference docs."""
    n, d = 10, 4
    c = _example_block(tc.Circuit(n), np.ones([2 * d, n]), d)
    nodes, _ = c._copy()
    (inp, out, sd), _ = cons.get_tn_info(nodes)
    path = paths.greedy(inp, out, sd)
    cst = paths.path_cost(inp, out, sd, path)
    k = KATS["greedy_cost_example_block"]
    # opt_einsum/cotengra of that era count 2 flops per inner-product MAC
    assert abs(np.log10(2 * cst["flops"]) - k["log10_flops"]) < 1e-3
    assert np.log2(cst["size"]) == k["log2_size"]
    assert abs(np.log2(cst["write"]) - k["log2_write"]) < 1e-3


def test_node_capture_counts():
    k = KATS["node_counts"]
    with cons.runtime_nodes_capture() as captured:
        c = tc.Circuit(3)
        c.h(0)
        c.amplitude("010")
    assert len(captured["nodes"]) == k["amplitude_capture"]
    with cons.runtime_nodes_capture() as captured:
        c = tc.Circuit(3)
        c.h(0)
        c.expectation_ps(z=[-3], reuse=False)
    assert len(captured["nodes"]) == k["expectation_capture"]


def test_lightcone_counts_and_value():
    def construct_c(pbc=True):
        n = 4
        ns = n if pbc else n - 1
        c = tc.Circuit(n)
        for j in range(2):
            for i in range(n):
                c.rx(i, theta=0.2, name="rx" + str(j) + "-" + str(i))
            for i in range(ns):
                c.cnot(i, (i + 1) % n, name="cnot" + str(j) + "-" + str(i))
        return c

    for b in [True, False]:
        c = construct_c(b)
        m1 = c.expectation_ps(z=[0], enable_lightcone=True)
        m2 = c.expectation_ps(z=[0])
        np.testing.assert_allclose(m1, m2, atol=1e-5)
        nodes = c.expectation_before([gates.z(), 0], reuse=False)
        l1 = len(nodes)
        nodes = ocirc._full_light_cone_cancel(nodes)
        l2 = len(nodes)
        want = KATS["lightcone_counts"]["pbc" if b else "open"]
        assert [l1, l2] == want


def test_merge_single_gates_equivalence():
    def build():
        c = tc.Circuit(6)
        for i in range(6):
            c.h(i)
        for i in range(5):
            c.cnot(i, i + 1)
            c.rx(i, theta=0.3 * (i + 1))
            c.rzz(i, i + 1, theta=0.11 * (i + 1))
        for i in range(6):
            c.ry(i, theta=0.7)
        return c

    with cons.runtime_contractor("greedy"):
        expected = build().state()
    with cons.runtime_contractor("greedy", preprocessing=True):
        got = build().state()
    np.testing.assert_allclose(got, expected, rtol=1e-5, atol=1e-5)
    with cons.runtime_contractor("plain"):
        got = build().state()
    np.testing.assert_allclose(got, expected, rtol=1e-5, atol=1e-5)


def test_edge_order_errors():
    c = tc.Circuit(2)
    c.h(0)
    nodes, d_edges = c._copy()
    with pytest.raises(ValueError, match="more than one remaining edge"):
        cons.contractor(nodes)
    nodes, d_edges = c._copy()
    with pytest.raises(ValueError, match="output edges are not equal"):
        cons.contractor(nodes, output_edge_order=d_edges[:1])


def test_diagonal_gate_equals_dense():
    # tests/test_hyperedge.py:530-559: c.diagonal == dense any(diagflat(d))
    d = np.exp(1j * np.arange(4) * 0.3)
    c1 = tc.Circuit(3)
    c1.h(0); c1.h(1); c1.h(2)
    c1.diagonal(0, 2, diag=d)
    c1.rx(1, theta=0.4)
    c2 = tc.Circuit(3)
    c2.h(0); c2.h(1); c2.h(2)
    c2.any(0, 2, unitary=np.diagflat(d))
    c2.rx(1, theta=0.4)
    np.testing.assert_allclose(c1.state(), c2.state(), atol=1e-5)
    np.testing.assert_allclose(c1.expectation_ps(z=[0], y=[1]), c2.expectation_ps(z=[0], y=[1]), atol=1e-5)
