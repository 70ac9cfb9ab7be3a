"""GPU tier: every C-ABI kernel on its own against numpy (the oracle's arithmetic), including edge shapes."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand_c(rng, shape):
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)


def _apply_np(state, n, qubits, m):
    k = len(qubits)
    psi = np.tensordot(m.reshape([2] * (2 * k)), state.reshape([2] * n), axes=[list(range(k, 2 * k)), list(qubits)])
    return np.ascontiguousarray(np.moveaxis(psi, list(range(k)), list(qubits))).reshape(-1)


@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (3, 1), (6, 3), (9, 5), (12, 2), (15, 4), (18, 1)])
def test_apply_dense(cuda, n, k):
    from tensorcircuit_ng_b200 import _lib

    rng = np.random.default_rng(n * 10 + k)
    psi = _rand_c(rng, 2**n)
    m = _rand_c(rng, (2**k, 2**k))
    qubits = [int(q) for q in rng.permutation(n)[:k]]
    st = torch.from_numpy(psi).cuda()
    mt = torch.from_numpy(m).cuda()
    _lib.call("tcb_sv_apply_dense", st.data_ptr(), n, 1, _lib.int_array([n - 1 - q for q in qubits]), k, mt.data_ptr(), 0,
              _lib.stream_ptr())  # fmt: skip
    want = _apply_np(psi, n, qubits, m)
    assert np.abs(st.cpu().numpy() - want).max() <= 1e-5 * max(1, np.abs(want).max())


@pytest.mark.parametrize("n,k,packed", [(1, 1, True), (4, 2, False), (10, 3, True), (13, 2, False), (16, 6, True)])
def test_apply_diag(cuda, n, k, packed):
    from tensorcircuit_ng_b200 import _lib

    rng = np.random.default_rng(n + k)
    psi = _rand_c(rng, 2**n)
    d = _rand_c(rng, 2**k)
    qubits = [int(q) for q in rng.permutation(n)[:k]]
    st = torch.from_numpy(psi).cuda()
    buf = torch.from_numpy(d if packed else np.diag(d).astype(np.complex64).copy()).cuda()
    stride = 1 if packed else 2**k + 1
    _lib.call("tcb_sv_apply_diag", st.data_ptr(), n, 1, _lib.int_array([n - 1 - q for q in qubits]), k, buf.data_ptr(),
              stride, 0, 0, _lib.stream_ptr())  # fmt: skip
    want = _apply_np(psi, n, qubits, np.diag(d))
    assert np.abs(st.cpu().numpy() - want).max() <= 1e-5 * max(1, np.abs(want).max())


def test_batched_dense_and_init(cuda):
    from tensorcircuit_ng_b200 import _lib

    n, B = 9, 5
    rng = np.random.default_rng(0)
    st = torch.empty(B * 2**n, dtype=torch.complex64, device="cuda")
    _lib.call("tcb_sv_init_zero", st.data_ptr(), n, B, _lib.stream_ptr())
    ref = np.zeros((B, 2**n), dtype=np.complex64)
    ref[:, 0] = 1
    assert np.array_equal(st.cpu().numpy().reshape(B, -1), ref)
    ms = _rand_c(rng, (B, 2, 2))
    mt = torch.from_numpy(ms).cuda()
    _lib.call("tcb_sv_apply_dense", st.data_ptr(), n, B, _lib.int_array([n - 1 - 3]), 1, mt.data_ptr(), 4, _lib.stream_ptr())
    got = st.cpu().numpy().reshape(B, -1)
    for b in range(B):
        assert np.abs(got[b] - _apply_np(ref[b], n, [3], ms[b])).max() <= 1e-6


@pytest.mark.parametrize("n", [1, 2, 7, 12, 17, 21])
def test_expectation_kernels(cuda, n):
    from tensorcircuit_ng_b200 import expect

    rng = np.random.default_rng(n)
    psi = _rand_c(rng, 2**n)
    psi /= np.linalg.norm(psi)
    st = torch.from_numpy(psi).cuda()
    p = np.abs(psi.astype(np.complex128)) ** 2
    idx = np.arange(2**n)
    # (n = 17 asks for more terms than one launch of the kernel carries: exercises the chunking)
    nterms = 400 if n == 17 else min(37, 3 * n)
    terms = [[int(q) for q in rng.permutation(n)[: int(rng.integers(1, min(n, 4) + 1))]] for _ in range(nterms)]
    got = expect.z_expectations(st, n, terms).cpu().numpy()
    for t, g in zip(terms, got):
        sign = np.ones(2**n)
        for q in t:
            sign *= 1 - 2 * ((idx >> (n - 1 - q)) & 1)
        assert abs(g - float(np.sum(p * sign))) <= 1e-6
    X = np.array([[0, 1], [1, 0]], dtype=np.complex64)
    Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex64)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex64)
    for trial in range(4):
        sites = [int(q) for q in rng.permutation(n)[: min(n, 3)]]
        kinds = [int(rng.integers(0, 3)) for _ in sites]
        phi = psi.copy()
        for q, kd in zip(sites, kinds):
            phi = _apply_np(phi, n, [q], [X, Y, Z][kd])
        want = np.vdot(psi.astype(np.complex128), phi.astype(np.complex128))
        xs = [q for q, kd in zip(sites, kinds) if kd == 0]
        ys = [q for q, kd in zip(sites, kinds) if kd == 1]
        zs = [q for q, kd in zip(sites, kinds) if kd == 2]
        got = complex(expect.pauli_expectation(st, n, xs, ys, zs).cpu().numpy())
        assert abs(got - want) <= 1e-6, (sites, kinds)


def _einsum_case(rng, modes_a, modes_b, modes_out):
    a = np.asarray(_rand_c(rng, [2] * len(modes_a)))
    b = np.asarray(_rand_c(rng, [2] * len(modes_b)))
    want = np.einsum(f"{modes_a},{modes_b}->{modes_out}", a, b)
    return a, b, want


@pytest.mark.parametrize(
    "ma,mb,mo",
    [("a", "ba", "b"), ("b", "b", ""), ("ab", "ba", ""), ("a", "b", "ab"), ("abc", "cd", "abd"), ("abc", "cd", "dba"),
     ("abcdefgh", "hgij", "abcdefij"), ("abcdefghij", "jihgklmn", "nmlkabcdef"), ("abx", "xcd", "cadb"),
     ("zab", "zbc", "zac"), ("abcdefghijkl", "lkjmno", "omnabcdefghi"), ("", "ab", "ba")],
)  # fmt: skip
def test_tn_contract_vs_einsum(cuda, ma, mb, mo):
    from tensorcircuit_ng_b200 import tnengine

    rng = np.random.default_rng(len(ma) * 31 + len(mb))
    a, b, want = _einsum_case(rng, ma, mb, mo)
    got = tnengine.contract_raw(torch.from_numpy(a).cuda(), list(ma), torch.from_numpy(b).cuda(), list(mb), list(mo))
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


def test_tn_contract_views_conj_and_grad(cuda):
    from tensorcircuit_ng_b200 import tnengine

    rng = np.random.default_rng(5)
    a = torch.from_numpy(_rand_c(rng, [2] * 6)).cuda()
    b = torch.from_numpy(_rand_c(rng, [2] * 5)).cuda()
    # permuted view + lazy conjugate + sliced (select) view, all consumed in place
    av = a.permute(3, 0, 5, 1, 4, 2)
    bv = b.conj().select(1, 1)
    got = tnengine.contract(av, list("abcdef"), bv, list("fegh"), list("hgabcd"))
    want = torch.einsum("abcdef,fegh->hgabcd", av, bv.resolve_conj())
    assert (got - want).abs().max() <= 2e-5
    # autograd of the pairwise contraction == torch's own einsum gradient
    a1 = a.clone().requires_grad_(True)
    b1 = b.clone().requires_grad_(True)
    out = tnengine.contract(a1, list("abcdef"), b1, list("fexyz"), list("zyabcdx"))
    w = torch.from_numpy(_rand_c(rng, [2] * 7)).cuda()
    (out * w).sum().real.backward()
    a2 = a.clone().requires_grad_(True)
    b2 = b.clone().requires_grad_(True)
    (torch.einsum("abcdef,fexyz->zyabcdx", a2, b2) * w).sum().real.backward()
    assert (a1.grad - a2.grad).abs().max() <= 1e-4 and (b1.grad - b2.grad).abs().max() <= 1e-4


def test_gate_grad_and_pack(cuda):
    from tensorcircuit_ng_b200 import _lib

    n = 12
    rng = np.random.default_rng(2)
    lam, psi = _rand_c(rng, 2**n), _rand_c(rng, 2**n)
    for qubits in ([3], [0, 7], [11, 2]):
        k = len(qubits)
        g = torch.zeros(4**k * 2, dtype=torch.float64, device="cuda")
        lt, pt = torch.from_numpy(lam).cuda(), torch.from_numpy(psi).cuda()
        _lib.call("tcb_sv_gate_grad", lt.data_ptr(), pt.data_ptr(), n, 1, _lib.int_array([n - 1 - q for q in qubits]), k,
                  g.data_ptr(), 0, _lib.stream_ptr())  # fmt: skip
        got = torch.view_as_complex(g.reshape(-1, 2)).cpu().numpy().reshape(2**k, 2**k)
        L = np.moveaxis(lam.reshape([2] * n), qubits, list(range(k))).reshape(2**k, -1).astype(np.complex128)
        P = np.moveaxis(psi.reshape([2] * n), qubits, list(range(k))).reshape(2**k, -1).astype(np.complex128)
        want = L @ P.conj().T
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max() + 1e-6
    st = torch.from_numpy(psi).cuda()
    buf = torch.empty(2 ** (n - 1), dtype=torch.complex64, device="cuda")
    for bit, want_v in [(0, 1), (5, 0), (n - 1, 1)]:
        _lib.call("tcb_sv_pack_half", st.data_ptr(), buf.data_ptr(), n, bit, want_v, _lib.stream_ptr())
        idx = np.arange(2**n)
        sel = psi[((idx >> bit) & 1) == want_v]
        assert np.array_equal(buf.cpu().numpy(), sel)
        st2 = torch.zeros_like(st)
        _lib.call("tcb_sv_unpack_half", st2.data_ptr(), buf.data_ptr(), n, bit, want_v, _lib.stream_ptr())
        ref = np.where(((idx >> bit) & 1) == want_v, psi, 0)
        assert np.array_equal(st2.cpu().numpy(), ref)


@pytest.mark.parametrize("n,qubits", [(1, [0]), (2, [1, 0]), (7, [3]), (12, [11]), (12, [0, 7]), (13, [12, 2]), (16, [5, 6]),
                                      (20, [19, 0])])  # fmt: skip
def test_adjoint_step_fused(cuda, n, qubits):
    """tcb_sv_adjoint_step == apply(U^dagger, psi); gate_grad(lam, psi_in); apply(U^dagger, lam), batch 1 and 3."""
    from tensorcircuit_ng_b200 import _lib

    k = len(qubits)
    rng = np.random.default_rng(n * 7 + k)
    for batch in (1, 3):
        lam, psi = _rand_c(rng, (batch, 2**n)), _rand_c(rng, (batch, 2**n))
        ud = _rand_c(rng, (batch, 2**k, 2**k))
        lt, pt, ut = torch.from_numpy(lam).cuda(), torch.from_numpy(psi).cuda(), torch.from_numpy(ud).cuda()
        g = torch.zeros(batch * 4**k * 2, dtype=torch.float64, device="cuda")
        _lib.call("tcb_sv_adjoint_step", lt.data_ptr(), pt.data_ptr(), n, batch, _lib.int_array([n - 1 - q for q in qubits]),
                  k, ut.data_ptr(), 4**k, g.data_ptr(), 4**k, _lib.stream_ptr())  # fmt: skip
        got_g = torch.view_as_complex(g.reshape(-1, 2)).cpu().numpy().reshape(batch, 2**k, 2**k)
        for b in range(batch):
            psi_in = _apply_np(psi[b].astype(np.complex128), n, qubits, ud[b].astype(np.complex128))
            lam_in = _apply_np(lam[b].astype(np.complex128), n, qubits, ud[b].astype(np.complex128))
            L = np.moveaxis(lam[b].reshape([2] * n), qubits, list(range(k))).reshape(2**k, -1).astype(np.complex128)
            P = np.moveaxis(psi_in.reshape([2] * n), qubits, list(range(k))).reshape(2**k, -1)
            want_g = L @ P.conj().T
            assert np.abs(pt[b].cpu().numpy() - psi_in).max() <= 2e-6 * np.abs(psi_in).max()
            assert np.abs(lt[b].cpu().numpy() - lam_in).max() <= 2e-6 * np.abs(lam_in).max()
            assert np.abs(got_g[b] - want_g).max() <= 2e-6 * np.abs(want_g).max() + 1e-6 * np.sqrt(2.0**n)


@pytest.mark.parametrize("n", [1, 2, 6, 13, 18])
def test_cross_marginals(cuda, n):
    """tcb_sv_cross_marginals: per gate, the marginal of lam * conj(psi) over the gate's bits (>= 2 launches' worth)."""
    from tensorcircuit_ng_b200 import _lib

    rng = np.random.default_rng(n)
    lam, psi = _rand_c(rng, 2**n), _rand_c(rng, 2**n)
    gates = []
    for _ in range(11):
        if n >= 2 and rng.integers(0, 2):
            a, b = (int(x) for x in rng.permutation(n)[:2])
            gates.append((a, b))
        else:
            gates.append((int(rng.integers(0, n)), -1))
    out = torch.zeros(len(gates) * 4, 2, dtype=torch.float64, device="cuda")
    lt, pt = torch.from_numpy(lam).cuda(), torch.from_numpy(psi).cuda()
    _lib.call("tcb_sv_cross_marginals", lt.data_ptr(), pt.data_ptr(), n, 1, len(gates),
              _lib.int_array([x for g in gates for x in g]), out.data_ptr(), 0, _lib.stream_ptr())  # fmt: skip
    got = torch.view_as_complex(out).cpu().numpy().reshape(len(gates), 4)
    q = lam.astype(np.complex128) * np.conj(psi.astype(np.complex128))
    idx = np.arange(2**n)
    for j, (a, b) in enumerate(gates):
        c = (idx >> a) & 1 if b < 0 else (((idx >> a) & 1) << 1) | ((idx >> b) & 1)
        for cc in range(4):
            want = q[c == cc].sum()
            assert abs(got[j, cc] - want) <= 2e-6 * np.abs(q).sum() ** 0.5 + 1e-6 * abs(want), (j, cc)


@pytest.mark.parametrize("n,sel", [(1, []), (2, []), (3, []), (5, [3, 4]), (12, [3, 5, 6, 8, 9, 10, 11]), (15, [4, 14]),
                                   (19, [7, 9, 11, 13, 15, 17, 18])])  # fmt: skip
def test_cross_rdm(cuda, n, sel):
    """tcb_sv_cross_rdm: out[t][r][c] = sum_rest lam[rest, bit_t = r] conj(psi[rest, bit_t = c]) for the low bits + sel."""
    from tensorcircuit_ng_b200 import _lib

    rng = np.random.default_rng(n + len(sel))
    lam, psi = _rand_c(rng, 2**n), _rand_c(rng, 2**n)
    low = min(3, n)
    bits = list(range(low)) + sel
    out = torch.zeros(40, 2, dtype=torch.float64, device="cuda")
    lt, pt = torch.from_numpy(lam).cuda(), torch.from_numpy(psi).cuda()
    _lib.call("tcb_sv_cross_rdm", lt.data_ptr(), pt.data_ptr(), n, 1, len(sel), _lib.int_array(sel) if sel else None,
              0, out.data_ptr(), 0, _lib.stream_ptr())  # fmt: skip
    if sel:  # skip_low: only the selected bits are written
        out2 = torch.zeros(40, 2, dtype=torch.float64, device="cuda")
        _lib.call("tcb_sv_cross_rdm", lt.data_ptr(), pt.data_ptr(), n, 1, len(sel), _lib.int_array(sel), 1,
                  out2.data_ptr(), 0, _lib.stream_ptr())  # fmt: skip
        assert float(out2[: min(3, n) * 4].abs().max()) == 0.0
        assert float((out2[min(3, n) * 4 :] - out[min(3, n) * 4 :]).abs().max()) <= 1e-9 * float(out.abs().max())
    got = torch.view_as_complex(out).cpu().numpy().reshape(10, 2, 2)
    L = lam.astype(np.complex128).reshape([2] * n)
    P = psi.astype(np.complex128).reshape([2] * n)
    for t, b in enumerate(bits):
        ax = n - 1 - b  # flat bit b is axis n-1-b
        Lm = np.moveaxis(L, ax, 0).reshape(2, -1)
        Pm = np.moveaxis(P, ax, 0).reshape(2, -1)
        want = Lm @ Pm.conj().T
        assert np.abs(got[t] - want).max() <= 2e-6 * np.abs(want).max() + 1e-6 * np.sqrt(2.0**n), (t, b)
    assert len(bits) == 10 or np.abs(got[len(bits):]).max() == 0


def test_error_reporting(cuda):
    from tensorcircuit_ng_b200 import _lib

    st = torch.zeros(8, dtype=torch.complex64, device="cuda")
    with pytest.raises(_lib.EngineError, match="duplicate bit"):
        _lib.call("tcb_sv_apply_dense", st.data_ptr(), 3, 1, _lib.int_array([1, 1]), 2, st.data_ptr(), 0, _lib.stream_ptr())
    with pytest.raises(_lib.EngineError, match="null pointer"):
        _lib.call("tcb_sv_init_zero", None, 3, 1, _lib.stream_ptr())
    with pytest.raises(_lib.EngineError, match="tile_bits"):
        _lib.call("tcb_sv_run_pass", st.data_ptr(), 3, 1, st.data_ptr(), 100, 13, 4, 0, st.data_ptr(), 0, 0, _lib.stream_ptr())
