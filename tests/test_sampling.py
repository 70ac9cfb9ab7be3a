"""SURVEY §8f rank 2 — sampling from the resident state.  CPU tier: the oracle restatement on closed-form
cases.  GPU tier: `tcb_sv_sample_prepare` / `tcb_sv_sample` through `Circuit.sample / measure` against it."""
import numpy as np
import pytest

import tc_oracle as otc
from tc_oracle import quantum as oq


def test_oracle_probability_sample_closed_forms():
    # Bell state over 3 qubits (tests/test_circuit.py:1462-1498): only |000> and |110>
    c = otc.Circuit(3)
    c.H(0)
    c.cnot(0, 1)
    p = np.abs(c.wavefunction()) ** 2
    got = oq.probability_sample(6, p, [0.0, 0.2, 0.49, 0.51, 0.8, 0.999])
    assert got.tolist() == [6, 6, 6, 0, 0, 0]  # r = 1 - u: small u -> the upper outcome
    bits, prob = oq.measure(c.wavefunction(), [0, 1, 2], [0.7, 0.3, 0.9])
    assert bits.tolist() == [1.0, 1.0, 0.0] and abs(prob - 0.5) < 1e-6  # u0 > 1/2 -> 1; then forced 1; forced 0
    bits, prob = oq.measure(c.wavefunction(), [0, 1, 2], [0.2, 0.99, 0.99])
    assert bits.tolist() == [0.0, 0.0, 0.0] and abs(prob - 0.5) < 1e-6
    assert oq.sample_int2bin(np.array([6, 1]), 3).tolist() == [[1, 1, 0], [0, 0, 1]]


def _random_circuit(mod, n, seed):
    rng = np.random.default_rng(seed)
    c = mod.Circuit(n)
    for q in range(n):
        c.h(q)
    for _ in range(3):
        for q in range(n):
            c.rx(q, theta=float(rng.uniform(0, 2 * np.pi)))
            c.rz(q, theta=float(rng.uniform(0, 2 * np.pi)))
        for q in range(n - 1):
            c.cnot(q, q + 1)
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 5, 11, 12, 13, 17, 20])
def test_gpu_cdf_sampling_matches_oracle(cuda, n):
    import torch

    import tensorcircuit_ng_b200 as tc

    c = _random_circuit(tc, n, n)
    psi = _random_circuit(otc, n, n).wavefunction()
    p = np.abs(psi.astype(np.complex128)) ** 2
    p /= p.sum()
    cuml = np.cumsum(p)
    rng = np.random.default_rng(100 + n)
    shots = 257
    u = rng.uniform(size=shots)
    u[:3] = [0.0, 0.5, 1.0 - 1e-12]
    got = c.sample(batch=shots, allow_state=True, status=torch.from_numpy(u), format="sample_int").cpu().numpy()
    want = oq.probability_sample(shots, p, u)
    r = cuml[-1] * (1.0 - u)
    # exact except where r sits within float32 state rounding of a CDF step: then a neighbour is also right
    tol = 3e-6
    for g, w, rr in zip(got, want, r):
        if g != w:
            lo = cuml[g - 1] if g > 0 else 0.0
            assert lo - tol <= rr <= cuml[g] + tol, (n, g, w, rr)
    assert np.mean(got == want) > 0.98
    # format None: (configuration, probability) pairs
    pairs = c.sample(batch=4, allow_state=True, status=torch.from_numpy(u[3:7]))
    for (conf, pr), w in zip(pairs, want[3:7]):
        assert conf.cpu().tolist() == oq.sample_int2bin(np.array(w), n).tolist()
        assert abs(float(pr) - p[w]) < 1e-6 + 1e-4 * p[w]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 4, 9, 12, 14, 18])
def test_gpu_perfect_sampling_matches_oracle_rule(cuda, n):
    import torch

    import tensorcircuit_ng_b200 as tc

    c = _random_circuit(tc, n, 50 + n)
    psi = _random_circuit(otc, n, 50 + n).wavefunction()
    rng = np.random.default_rng(7 * n)
    shots = 24
    u = rng.uniform(size=(shots, n))
    got = c.sample(batch=shots, allow_state=False, status=torch.from_numpy(u), format="sample_bin").cpu().numpy()
    agree = 0
    for s in range(shots):
        bits, prob = oq.measure(psi, list(range(n)), u[s])
        agree += int(got[s].tolist() == bits.astype(int).tolist())
    assert agree >= shots - 1  # a uniform within 1e-6 of a conditional probability may flip one walk
    bits, prob = c.perfect_sampling(status=torch.from_numpy(u[0]))
    wb, wp = oq.measure(psi, list(range(n)), u[0])
    assert bits.cpu().tolist() == wb.tolist() and abs(float(prob) - wp) < 1e-6 + 1e-4 * wp
    if n >= 4:
        sub = [n - 1, 0, 2]
        bits, prob = c.measure(*sub, with_prob=True, status=torch.from_numpy(u[1, :3]))
        wb, wp = oq.measure(psi, sub, u[1, :3])
        assert bits.cpu().tolist() == wb.tolist() and abs(float(prob) - wp) < 1e-6 + 1e-4 * wp


@pytest.mark.gpu
def test_gpu_sample_formats_and_statistics(cuda):  # tests/test_circuit.py:1462-1498,:1646-1690
    import torch

    import tensorcircuit_ng_b200 as tc

    c = tc.Circuit(3)
    c.H(0)
    c.cnot(0, 1)
    assert len(c.sample()) == 2 and len(c.sample(allow_state=True)) == 2
    r = c.sample(batch=8, status=np.random.uniform(size=[8, 3]))
    assert len(r) == 8
    for conf, prob in r:
        assert len(conf) == 3 and 0.0 <= float(prob) <= 1.0
    g = torch.Generator(device="cpu").manual_seed(42)
    assert len(c.sample(batch=8, allow_state=True, random_generator=g)) == 8
    assert len(c.sample(batch=8, allow_state=True, status=np.random.uniform(size=[8]), format="sample_bin")) == 8
    c2 = tc.Circuit(2)
    c2.H(0)
    c2.cnot(0, 1)
    for allow_state in (False, True):
        for batch in (None, 1, 3):
            nb = 1 if batch is None else batch
            assert len(c2.sample(batch=batch, allow_state=allow_state, format="sample_int", random_generator=g)) == nb
            sb = c2.sample(batch=batch, allow_state=allow_state, format_="sample_bin", random_generator=g)
            assert len(sb) == nb and len(sb[0]) == 2
            assert len(c2.sample(batch=batch, allow_state=allow_state, format="count_vector", random_generator=g)) == 4
            sm, ct = c2.sample(batch=batch, allow_state=allow_state, format="count_tuple", random_generator=g)
            assert len(sm) == len(ct) >= 1
            for k in c2.sample(batch=batch, allow_state=allow_state, format="count_dict_bin", random_generator=g):
                assert len(k) == 2 and k in ("00", "11")
            for k in c2.sample(batch=batch, allow_state=allow_state, format="count_dict_int", random_generator=g):
                assert k in (0, 3)
    # statistics: 20000 shots of a 10-qubit state reproduce its distribution (total variation < 5 %)
    n = 10
    c3 = _random_circuit(tc, n, 3)
    p = c3.probability().double().cpu().numpy()
    for allow_state in (True, False):
        cv = c3.sample(batch=20000, allow_state=allow_state, format="count_vector", random_generator=g).cpu().numpy()
        assert cv.sum() == 20000
        assert 0.5 * np.abs(cv / 20000.0 - p / p.sum()).sum() < 0.12


def test_sample_format_helpers_on_host_tensors():
    """quantum.sample2all and friends (tensorcircuit/quantum.py:3587-3902) on explicit samples."""
    import torch

    from tensorcircuit_ng_b200 import quantum as q

    s = torch.tensor([0, 3, 3, 2])
    assert q.sample_int2bin(s, 2).tolist() == [[0, 0], [1, 1], [1, 1], [1, 0]]
    assert q.sample_int2bin(s, 2).tolist() == oq.sample_int2bin(s.numpy(), 2).tolist()
    assert q.sample_bin2int(q.sample_int2bin(s, 2), 2).tolist() == s.tolist()
    assert q.sample2all(s, 2, format="sample_int").tolist() == s.tolist()
    assert q.sample2all(q.sample_int2bin(s, 2), 2, format="sample_int").tolist() == s.tolist()
    assert q.sample2all(s, 2, format="count_vector").tolist() == [1, 0, 1, 2]
    idx, cnt = q.sample2all(s, 2, format="count_tuple")
    assert idx.tolist() == [0, 2, 3] and cnt.tolist() == [1, 1, 2]
    assert q.sample2all(s, 2, format="count_dict_bin") == {"00": 1, "10": 1, "11": 2}
    assert q.sample2all(s, 2, format="count_dict_int") == {0: 1, 2: 1, 3: 2}
    with pytest.raises(ValueError):
        q.sample2all(s, 2, format="nonsense")
    with pytest.raises(ValueError):
        q.sample2all(torch.zeros(2, 2, 2), 2)
