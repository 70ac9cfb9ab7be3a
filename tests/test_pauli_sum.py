"""SURVEY §8f rank 1 — Pauli-string-sum operators.  CPU tier: the oracle restatement against the reference's
own tests / docstring values, and the engine's host tables against the oracle.  GPU tier: `tcb_sv_pauli_sum`
(through the public API) against the oracle, including gradients."""
import itertools

import numpy as np
import pytest

import tc_oracle as otc
from tc_oracle import quantum as oq

# the Hamiltonian of the reference's tests/test_quantum.py:1534-1545
REF_LS = [[1, 0, 0, 0], [0, 0, 3, 0], [3, 3, 0, 0], [1, 2, 0, 0], [0, 2, 2, 3]]
REF_W = [0.5, -0.3, 1.2, 0.7, -0.9]


def _rand_terms(rng, n, nterms, complex_w=False):
    ls = rng.integers(0, 4, size=(nterms, n)).tolist()
    w = rng.normal(size=nterms)
    if complex_w:
        w = w + 1j * rng.normal(size=nterms)
    return ls, w.tolist()


# ------------------------------------------------------------------ oracle vs the reference's values
def test_oracle_mvp_matches_dense():  # tests/test_quantum.py:1534-1576
    rng = np.random.default_rng(0)
    psi = rng.normal(size=16) + 1j * rng.normal(size=16)
    mvp = oq.PauliStringSum2MVP(REF_LS, REF_W)
    dense = oq.PauliStringSum2Dense(REF_LS, REF_W)
    np.testing.assert_allclose(mvp(psi), dense @ psi, atol=1e-12)
    np.testing.assert_allclose(mvp(psi.reshape((2,) * 4)).reshape(-1), dense @ psi, atol=1e-12)
    np.testing.assert_allclose(oq.PauliStringSum2MVP([], [])(psi), np.zeros_like(psi))
    psi3 = np.array([1.0, 0, 0, 0, 0, 0, 0, 1.0]) + 0j
    np.testing.assert_allclose(oq.PauliStringSum2MVP([[0, 0, 0]], [2.0])(psi3), 2.0 * psi3)


def test_oracle_heisenberg_spectrum():  # docstring of tensorcircuit/quantum.py:2148-2153 (Line1D(6), pbc)
    edges = [(i, (i + 1) % 6) for i in range(6)]
    ls, ws = oq.heisenberg_hamiltonian_terms(edges, 6)
    ev = np.linalg.eigvalsh(oq.PauliStringSum2Dense(ls, ws))
    np.testing.assert_allclose(ev[:6], [-11.2111025, -8.4721365, -8.472136, -8.472136, -6.0, -5.123106], atol=2e-5)


def test_oracle_operator_expectation_kat():  # tests/test_templates.py:190-211: 0.84147, gradient 0.54032
    h = oq.PauliStringSum2Dense([[1, 0]])

    def f(theta):
        c = otc.Circuit(2)
        c.ry(0, theta=theta)
        c.H(1)
        return oq.operator_expectation(c, h)

    assert abs(f(1.0) - 0.84147) < 1e-4
    assert abs((f(1.0 + 1e-3) - f(1.0 - 1e-3)) / 2e-3 - 0.54032) < 1e-3


def test_oracle_u1_sum_z():  # tests/test_quantum.py:1437-1447: <sum Z> = 8 - 2 i on Hamming-weight-i states
    n = 8
    ls, ws = oq.heisenberg_hamiltonian_terms([(i, (i + 1) % n) for i in range(n)], n, hzz=0, hxx=0, hyy=0, hz=1)
    h = oq.PauliStringSum2Dense(ls, ws)
    weight = np.array([bin(i).count("1") for i in range(1 << n)])
    for i in range(n + 1):
        s = (weight == i).astype(np.complex128)
        s /= np.linalg.norm(s)
        assert abs(np.real(np.vdot(s, h @ s)) - (n - 2 * i)) < 1e-9


# ------------------------------------------------------------------ engine host tables vs the oracle (CPU)
@pytest.mark.parametrize("seed", range(4))
def test_host_tables_match_oracle_dense(seed):
    from tensorcircuit_ng_b200 import quantum as q

    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 7))
    ls, w = _rand_terms(rng, n, int(rng.integers(1, 40)), complex_w=bool(seed % 2))
    h = q.PauliStringSum(ls, w)
    np.testing.assert_allclose(h.to_dense_numpy(), oq.PauliStringSum2Dense(ls, w), atol=1e-5)
    assert h.hermitian == (seed % 2 == 0)
    assert np.all(np.diff(h.xmask.astype(np.int64)) >= 0)  # sorted by flip mask: one state read per run
    np.testing.assert_allclose(h.adjoint().to_dense_numpy(), oq.PauliStringSum2Dense(ls, w).conj().T, atol=1e-5)


def test_host_merges_duplicates_and_drops_zeros():
    from tensorcircuit_ng_b200 import quantum as q

    h = q.PauliStringSum([[1, 3], [1, 3], [0, 0], [2, 2]], [0.5, -0.5, 0.0, 1.0])
    assert h.nterms == 1 and h.xmask[0] == 3 and h.zmask[0] == 3
    assert abs(h.coef[0] - (-1.0)) < 1e-7  # YY carries i^2
    with pytest.raises(ValueError):
        q.PauliStringSum([[4, 0]])
    with pytest.raises(ValueError):
        q.PauliStringSum([])
    g = __import__("tensorcircuit_ng_b200").templates.graphs.Line1D(6)
    hh = q.heisenberg_hamiltonian(g)
    np.testing.assert_allclose(np.linalg.eigvalsh(hh.to_dense_numpy())[:2], [-11.2111025, -8.4721365], atol=2e-5)


# ------------------------------------------------------------------ GPU: the kernel through the public API
@pytest.mark.gpu
@pytest.mark.parametrize("n,nterms,cw", [(1, 3, False), (2, 7, True), (4, 5, False), (9, 60, True), (12, 200, False),
                                         (10, 2500, True), (15, 30, False)])  # fmt: skip
def test_gpu_mvp_and_expectation_vs_oracle(cuda, n, nterms, cw):
    import torch

    from tensorcircuit_ng_b200 import quantum as q

    rng = np.random.default_rng(n * 1000 + nterms)
    ls, w = (REF_LS, REF_W) if (n, nterms) == (4, 5) else _rand_terms(rng, n, nterms, complex_w=cw)
    psi = (rng.normal(size=2**n) + 1j * rng.normal(size=2**n)).astype(np.complex64)
    want = oq.PauliStringSum2MVP(ls, w)(psi.astype(np.complex128))
    h = q.PauliStringSum(ls, w)
    t = torch.from_numpy(psi).cuda()
    got = h.mvp(t).cpu().numpy()
    scale = np.abs(want).max() + 1e-30
    assert np.abs(got - want).max() <= 2e-5 * scale * max(1.0, np.sqrt(nterms) / 4)
    got_nd = q.PauliStringSum2MVP(ls, w)(t.reshape((2,) * n))
    assert tuple(got_nd.shape) == (2,) * n
    assert np.abs(got_nd.reshape(-1).cpu().numpy() - want).max() <= 2e-5 * scale * max(1.0, np.sqrt(nterms) / 4)
    e = complex(h.expectation(t).cpu())
    e_want = np.vdot(psi.astype(np.complex128), want)
    assert abs(e - e_want) <= 3e-5 * (abs(e_want) + np.linalg.norm(psi) ** 2 * scale / np.abs(psi).max())


@pytest.mark.gpu
def test_gpu_mvp_empty_and_identity(cuda):  # tests/test_quantum.py:1564-1576
    import torch

    from tensorcircuit_ng_b200 import quantum as q

    psi3 = torch.tensor([1.0, 0, 0, 0, 0, 0, 0, 1.0], dtype=torch.complex64).cuda()
    assert torch.equal(q.PauliStringSum2MVP([], [])(psi3), torch.zeros_like(psi3))
    np.testing.assert_allclose(q.PauliStringSum2MVP([[0, 0, 0]], [2.0])(psi3).cpu().numpy(), 2.0 * psi3.cpu().numpy(), atol=1e-6)
    z = q.PauliStringSum([[1, 1, 0], [1, 1, 0]], [1.0, -1.0])  # cancels to the empty sum
    assert z.nterms == 0 and float(z.mvp(psi3).abs().max()) == 0.0 and abs(complex(z.expectation(psi3).cpu())) == 0.0


@pytest.mark.gpu
def test_gpu_operator_expectation_kats(cuda):  # tests/test_templates.py:43-60 and :190-211
    import torch

    import tensorcircuit_ng_b200 as tc

    sparse = tc.quantum.PauliString2COO([1, 0])
    dense = torch.tensor(np.kron(np.array([[0, 1], [1, 0]]), np.eye(2)), dtype=torch.complex64)
    for h in (dense, sparse):
        def f(theta):
            c = tc.Circuit(2)
            c.ry(0, theta=theta)
            c.H(1)
            return tc.templates.measurements.operator_expectation(c, h)

        v, g = tc.backend.value_and_grad(f)(torch.ones([]))
        assert abs(float(v) - 0.84147) < 1e-4 and abs(float(g) - 0.54032) < 1e-4

    ham = tc.quantum.PauliString2COO([1])

    def f2(param):
        c = tc.Circuit(1)
        c.rx(0, theta=param[0])
        c.H(0)
        return tc.templates.measurements.sparse_expectation(c, ham)

    v, g = tc.backend.value_and_grad(f2)(torch.zeros([1]))
    assert abs(float(v) - 1.0) < 1e-4 and abs(float(g[0])) < 1e-4


@pytest.mark.gpu
def test_gpu_u1_sum_z(cuda):  # tests/test_quantum.py:1437-1447
    import torch

    import tensorcircuit_ng_b200 as tc

    n = 8
    sumz = tc.quantum.heisenberg_hamiltonian(tc.templates.graphs.Line1D(n), hzz=0, hxx=0, hyy=0, hz=1)
    weight = np.array([bin(i).count("1") for i in range(1 << n)])
    for i in range(n + 1):
        s = (weight == i).astype(np.complex64)
        s /= np.linalg.norm(s)
        c = tc.Circuit(n, inputs=torch.from_numpy(s).cuda())
        assert abs(float(tc.templates.measurements.operator_expectation(c, sumz)) - (n - 2 * i)) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("cw", [False, True])
def test_gpu_gradients_vs_dense_autograd(cuda, cw):
    import torch

    from tensorcircuit_ng_b200 import quantum as q

    rng = np.random.default_rng(5 + cw)
    n = 6
    ls, w = _rand_terms(rng, n, 25, complex_w=cw)
    h = q.PauliStringSum(ls, w)
    dense = torch.from_numpy(oq.PauliStringSum2Dense(ls, w).astype(np.complex64)).cuda()
    psi0 = torch.from_numpy((rng.normal(size=2**n) + 1j * rng.normal(size=2**n)).astype(np.complex64)).cuda()
    probe = torch.from_numpy((rng.normal(size=2**n) + 1j * rng.normal(size=2**n)).astype(np.complex64)).cuda()

    def losses(op_mvp, op_exp):
        a = psi0.clone().requires_grad_(True)
        l1 = (probe.conj() * op_mvp(a)).sum().real + (op_mvp(a).abs() ** 2).sum()
        (g1,) = torch.autograd.grad(l1, a)
        b = psi0.clone().requires_grad_(True)
        e = op_exp(b)
        l2 = 0.7 * e.real - 0.3 * e.imag
        (g2,) = torch.autograd.grad(l2, b)
        return g1, g2

    g1, g2 = losses(h.mvp, h.expectation)
    r1, r2 = losses(lambda v: dense @ v, lambda v: torch.vdot(v, dense @ v))
    assert float((g1 - r1).abs().max()) <= 2e-4 * float(r1.abs().max())
    assert float((g2 - r2).abs().max()) <= 2e-4 * float(r2.abs().max())


@pytest.mark.gpu
def test_gpu_tfim_energy_matches_per_term_expectation_ps(cuda):
    """The config-2 energy (examples/benchmark_jax_vs_torch_vqe.py:168-186) two ways: 2n-1 expectation_ps calls
    vs ONE Pauli-sum launch; values and parameter gradients agree, and both match the oracle."""
    import torch

    import tensorcircuit_ng_b200 as tc

    n, depth = 10, 2
    ls, ws = [], []
    for qb in range(n - 1):
        s = [0] * n
        s[qb] = s[qb + 1] = 3
        ls.append(s)
        ws.append(-1.0)
    for qb in range(n):
        s = [0] * n
        s[qb] = 1
        ls.append(s)
        ws.append(-1.0)
    ham = tc.quantum.PauliStringSum2COO(ls, ws)

    def ansatz(mod, p):
        c = mod.Circuit(n)
        for qb in range(n):
            c.h(qb)
        for l in range(depth):
            for qb in range(n - 1):
                c.rzz(qb, qb + 1, theta=p[l, 0, qb])
            for qb in range(n):
                c.rx(qb, theta=p[l, 1, qb])
        return c

    def e_terms(p):
        c = ansatz(tc, p)
        e = 0.0
        for qb in range(n - 1):
            e = e - c.expectation_ps(z=[qb, qb + 1]).real
        for qb in range(n):
            e = e - c.expectation_ps(x=[qb]).real
        return e

    def e_sum(p):
        return tc.templates.measurements.operator_expectation(ansatz(tc, p), ham)

    p = 0.3 * torch.randn(depth, 2, n, generator=torch.Generator(device="cpu").manual_seed(1), device="cpu").cuda()
    v1, g1 = tc.backend.value_and_grad(e_terms)(p)
    v2, g2 = tc.backend.value_and_grad(e_sum)(p)
    assert abs(float(v1) - float(v2)) < 2e-5 * n
    assert float((g1 - g2).abs().max()) < 5e-5
    want = oq.operator_expectation(ansatz(otc, p.cpu().numpy()), oq.PauliStringSum2Dense(ls, ws))
    assert abs(float(v2) - want) < 2e-5 * n
