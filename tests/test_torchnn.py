"""SURVEY §8f rank 4 — `TorchLayer` / `torch_interface` (tensorcircuit/torchnn.py:16-99, interfaces/torch.py:17-125):
the reference's tests/test_torchnn.py:21-49 on the engine, values and weight gradients checked against the oracle."""
import numpy as np
import pytest

import tc_oracle as otc

pytestmark = pytest.mark.gpu


def _qpred(mod, n, nlayers, stack, real):
    def qpred(x, weights):
        c = mod.Circuit(n)
        for i in range(n):
            c.rx(i, theta=x[i])
        for j in range(nlayers):
            for i in range(n - 1):
                c.cnot(i, i + 1)
            for i in range(n):
                c.rx(i, theta=weights[2 * j, i])
                c.ry(i, theta=weights[2 * j + 1, i])
        return real(stack([c.expectation_ps(x=[i]) for i in range(n)]))

    return qpred


def test_quantumnet_values_and_training_gradients(cuda):
    import torch

    import tensorcircuit_ng_b200 as tc

    n, nlayers = 6, 2
    ql = tc.TorchLayer(_qpred(tc, n, nlayers, tc.backend.stack, tc.backend.real), weights_shape=[2 * nlayers, n],
                       use_interface=False)  # fmt: skip
    yp = ql(torch.ones([3, n]))
    assert tuple(yp.shape) == (3, n)
    w = ql.q_weights[0].detach().cpu().numpy().astype(np.float64)
    oracle = _qpred(otc, n, nlayers, np.stack, np.real)
    want = oracle(np.ones(n), w)
    assert np.abs(yp.detach().cpu().numpy() - want[None, :]).max() < 2e-5
    # one optimiser step through the engine's vjps; gradient against the exact parameter-shift rule on the oracle
    x = torch.linspace(0.1, 0.9, n)
    loss = ql(x[None, :])[0].sum()
    loss.backward()
    g = ql.q_weights[0].grad.cpu().numpy()
    xh = x.cpu().numpy().astype(np.float64)
    f0 = lambda ww: float(np.sum(oracle(xh, ww)))
    from helpers import param_shift

    checks = [(0, 0), (1, 3), (3, 5), (2, 2), (0, 4), (3, 0)]
    ps = {c: param_shift(f0, w, c, "half") for c in checks}  # every weight feeds one rx / ry gate
    scale = max(abs(v) for v in ps.values())
    for c in checks:
        assert abs(g[c] - ps[c]) <= 1e-4 * max(abs(ps[c]), scale), (c, g[c], ps[c])
    torch.optim.SGD(ql.parameters(), lr=0.1).step()


def test_torch_interface_moves_host_tensors(cuda):
    import torch

    import tensorcircuit_ng_b200 as tc

    def f(params):
        c = tc.Circuit(1)
        c.rx(0, theta=params[0])
        c.ry(0, theta=params[1])
        return c.expectation([tc.gates.z(), [0]]).real

    f_torch = tc.interfaces.torch_interface(f, jit=True)
    a = torch.ones([2], requires_grad=True, device="cpu")
    b = f_torch(a)
    assert b.device.type == "cpu"
    (b**2).backward()
    want = np.cos(1.0) ** 2  # <Z> after rx(1) ry(1) on |0>
    assert abs(float(b) - want) < 1e-5
    assert a.grad is not None and tuple(a.grad.shape) == (2,) and bool(torch.isfinite(a.grad).all())
