"""Edge cases of the public API on the GPU path (empty circuits, tiny registers, `inputs=` states, wide networks that
only the tensor-network route can take, degenerate batches), each against the oracle or a closed form."""
import numpy as np
import pytest

import tc_oracle as otc

pytestmark = pytest.mark.gpu


def test_empty_and_tiny_circuits(cuda):
    import torch

    import tensorcircuit_ng_b200 as tc

    w = tc.Circuit(1).wavefunction().cpu().numpy()
    assert np.allclose(w, [1, 0])
    c = tc.Circuit(3)
    c.h(1)
    assert abs(float(c.expectation_ps(z=[0]).real) - 1.0) < 1e-6  # untouched qubit
    assert abs(float(c.expectation_ps(x=[1]).real) - 1.0) < 1e-6
    assert abs(float(c.expectation_ps(z=[1]).real)) < 1e-6
    c1 = tc.Circuit(1)
    c1.x(0)
    r = c1.sample(batch=4, allow_state=True, format="sample_int")
    assert r.cpu().tolist() == [1, 1, 1, 1]
    bits, p = c1.perfect_sampling(status=torch.tensor([0.3]))
    assert bits.cpu().tolist() == [1.0] and abs(float(p) - 1.0) < 1e-6
    assert abs(complex(c1.amplitude("1").cpu()) - 1.0) < 1e-6


def test_inputs_state_and_dense_operator(cuda):
    import torch

    import tensorcircuit_ng_b200 as tc

    rng = np.random.default_rng(0)
    n = 4
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi = (psi / np.linalg.norm(psi)).astype(np.complex64)
    c, oc = tc.Circuit(n, inputs=torch.from_numpy(psi).cuda()), otc.Circuit(n, inputs=psi)
    for cc in (c, oc):
        cc.cnot(0, 2)
        cc.rx(1, theta=0.4)
        cc.rzz(2, 3, theta=1.1)
    assert np.abs(c.wavefunction().cpu().numpy() - oc.wavefunction()).max() < 1e-6
    assert abs(complex(c.amplitude("0110").cpu()) - complex(oc.amplitude("0110"))) < 1e-6
    h = rng.normal(size=(2**n, 2**n)) + 1j * rng.normal(size=(2**n, 2**n))
    h = (h + h.conj().T).astype(np.complex64)
    got = float(tc.templates.measurements.operator_expectation(c, torch.from_numpy(h).cuda()))
    w = oc.wavefunction().astype(np.complex128)
    assert abs(got - float(np.real(np.vdot(w, h.astype(np.complex128) @ w)))) < 1e-4


def test_wide_register_goes_through_the_tree_route(cuda):
    """40 qubits: no statevector possible; amplitudes and local expectations of a GHZ-like circuit via the contraction tree."""
    import tensorcircuit_ng_b200 as tc

    n = 40
    c = tc.Circuit(n)
    c.h(0)
    for q in range(n - 1):
        c.cnot(q, q + 1)
    assert abs(complex(c.amplitude("0" * n).cpu()) - 2**-0.5) < 1e-6
    assert abs(complex(c.amplitude("1" * n).cpu()) - 2**-0.5) < 1e-6
    assert abs(complex(c.amplitude("0" * (n - 1) + "1").cpu())) < 1e-7


def test_degenerate_batches_and_constant_circuits(cuda):
    import torch

    import tensorcircuit_ng_b200 as tc

    def f(p):
        c = tc.Circuit(3)
        c.h(0)
        c.cnot(0, 1)
        c.rx(2, theta=p[0])
        return c.expectation_ps(z=[2]).real

    v, g = tc.backend.vvag(f, argnums=0, vectorized_argnums=0)(torch.tensor([[0.7]]))  # batch of one
    assert tuple(v.shape) == (1,) and abs(float(v[0]) - np.cos(0.7)) < 1e-6 and abs(float(g[0, 0]) + np.sin(0.7)) < 1e-5

    def const(p):  # nothing depends on p
        c = tc.Circuit(2)
        c.h(0)
        c.cnot(0, 1)
        return c.expectation_ps(z=[0, 1]).real + 0.0 * p.sum()

    v, g = tc.backend.value_and_grad(const)(torch.ones(3))
    assert abs(float(v) - 1.0) < 1e-6 and float(g.abs().max()) == 0.0


def test_reference_closed_forms(cuda):
    """tests/test_circuit.py:1501-1504 (<Y> = -1 on (-1, i)/sqrt 2), :57-62 (measure after a Toffoli reads a bit),
    :65-77 (Bell measurements are perfectly correlated)."""
    import torch

    import tensorcircuit_ng_b200 as tc

    c = tc.Circuit(1, inputs=torch.tensor(1 / np.sqrt(2) * np.array([-1, 1.0j]), dtype=torch.complex64).cuda())
    assert abs(complex(c.expectation_ps(y=[0]).cpu()) + 1.0) < 1e-5
    c = tc.Circuit(3)
    c.H(0)
    c.h(1)
    c.toffoli(0, 1, 2)
    assert float(c.measure(2)[0][0]) in (0.0, 1.0)
    g = torch.Generator(device="cpu").manual_seed(5)
    c = tc.Circuit(2)
    c.H(0)
    c.cnot(0, 1)
    counts = c.sample(batch=300, allow_state=False, format="count_dict_bin", random_generator=g)
    assert set(counts) <= {"00", "11"} and sum(counts.values()) == 300 and min(counts.get("00", 0), counts.get("11", 0)) > 90
    bits, p = c.measure(0, 1, with_prob=True, status=torch.tensor([0.9, 0.1]))
    assert bits.cpu().tolist() == [1.0, 1.0] and abs(float(p) - 0.5) < 1e-6


@pytest.mark.parametrize("n,with_prefix", [(12, True), (14, True), (15, False), (17, True)])
def test_generated_start_equals_init_kernel_plus_passes(cuda, n, with_prefix):
    """`tcb_sv_run_pass_generate` (the first pass builds its tiles of the initial product state in shared memory) against
    the init kernel + in-place passes (same to rounding), and both against the oracle: |0...0> and absorbed leading gates,
    the smallest state the T = 12 kernel takes and wider ones."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from helpers import build, oracle_state
    from tensorcircuit_ng_b200 import svengine

    rng = np.random.default_rng(n)
    ops = []
    if with_prefix:
        for q in range(n):
            ops.append(("h", [q], {}) if q % 3 else ("ry", [q], {"theta": float(rng.uniform(0, 6))}))
    for _ in range(2):
        for q in range(n - 1):
            ops.append(("rzz", [q, q + 1], {"theta": float(rng.uniform(0, 6))}))
        for q in range(n):
            ops.append(("rx", [q], {"theta": float(rng.uniform(0, 6))}))
        ops.append(("cnot", [0, n - 1], {}))
    out = {}
    for fuse in (True, False):
        svengine.fuse_start = fuse
        try:
            out[fuse] = build(tc, n, ops).wavefunction().cpu().numpy()
        finally:
            svengine.fuse_start = True
    assert np.abs(out[True] - out[False]).max() <= 3e-7  # (the product is associated differently: rounding only)
    assert np.abs(out[True] - oracle_state(n, ops)).max() <= 1e-5
