"""The plan objects of the C ABI (include/tcb200.h `tcb_sv_plan_*`, `tcb_tn_plan_*`; SURVEY §8b export list).
CPU tier: creation / validation / workspace packing of tensor-network plans (no device needed).
GPU tier: plan execution == the step-by-step executors == the oracle."""
import ctypes

import numpy as np
import pytest

import tc_oracle as otc


def _tree(rows, cols, depth, bits, target):
    import bench
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import planner
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    nodes_fn = lambda _: bench.build_rcs(tc, rows, cols, depth).amplitude_before(bits)  # noqa: E731
    inp, out, sd, tensors, groups = DistributedContractor._network(nodes_fn, None, True)
    td = planner.search_elimination(inp, out, sd, target_size=target, groups=groups)
    return td, tensors, sd


def test_tn_plan_create_workspace_and_validation(built):
    from tensorcircuit_ng_b200 import _lib, tnengine

    td, _, _ = _tree(3, 4, 8, "0" * 12, 2**4)
    assert len(td["sliced_inds"]) >= 1
    tp = tnengine.TreePlan(td["inputs"], td["output"], td["path"], list(td["sliced_inds"]))
    steps = tnengine.build_schedule(td["inputs"], td["output"], td["path"], sorted(td["sliced_inds"]))
    assert tp.nsteps == len(steps) == int(_lib.load().tcb_tn_plan_launches(tp._handle))
    assert int(_lib.load().tcb_tn_plan_output_elems(tp._handle)) == 1
    # liveness packing: the workspace is far smaller than the sum of all intermediates, and at least as large
    # as the biggest one
    td2, _, _ = _tree(4, 5, 8, "0" * 20, 2**12)
    tp2 = tnengine.TreePlan(td2["inputs"], td2["output"], td2["path"], list(td2["sliced_inds"]))
    steps2 = tnengine.build_schedule(td2["inputs"], td2["output"], td2["path"], sorted(td2["sliced_inds"]))
    sizes = [8 * 2 ** len(keep) for *_, keep, _ in steps2[:-1]]
    assert max(sizes) <= tp2.ws_bytes < 0.6 * sum(max(256, x) for x in sizes)
    # malformed SSA tables are rejected with a message, not executed
    lib = _lib.load()
    h = ctypes.c_void_p()
    leaf = (ctypes.c_int64 * 2)(2, 2)
    ids = (ctypes.c_int32 * 3)(0, 0, 2)  # a == b
    d = (_lib.ContractDesc * 1)()
    oe = (ctypes.c_int64 * 1)(1)
    rc = lib.tcb_tn_plan_create(2, leaf, 1, ids, d, oe, 0, None, None, ctypes.byref(h))
    assert rc != 0 and b"not a valid SSA step" in lib.tcb_last_error()
    ids = (ctypes.c_int32 * 3)(0, 1, 1)  # writes over a leaf
    assert lib.tcb_tn_plan_create(2, leaf, 1, ids, d, oe, 0, None, None, ctypes.byref(h)) != 0
    assert lib.tcb_tn_plan_workspace_size(None) == 0 and lib.tcb_tn_plan_destroy(None) == 0
    assert lib.tcb_sv_plan_workspace_size(None) == 0 and lib.tcb_sv_plan_destroy(None) == 0
    # a statevector plan with no device programs can be created and destroyed without a GPU
    gates = np.array([[1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [2, 4, 1, 0, 0, 0, 0, 0, 0, 0], [3, 20, 2, 1, 0, 0, 0, 0, 0, 0]],
                     dtype=np.int64)  # fmt: skip
    assert lib.tcb_sv_plan_create(3, None, 0, None, 0, gates.ctypes.data, 3, ctypes.byref(h)) == 0
    assert lib.tcb_sv_plan_launches(h, 1) == 4 and lib.tcb_sv_plan_launches(h, 0) == 0  # a 3-qubit gate: 2 launches
    assert lib.tcb_sv_plan_destroy(h) == 0
    bad = np.array([[8, 0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.int64)
    assert lib.tcb_sv_plan_create(3, None, 0, None, 0, bad.ctypes.data, 1, ctypes.byref(h)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,depth,target", [(3, 3, 6, 2**3), (3, 4, 8, 2**4), (4, 4, 8, 2**6), (4, 5, 6, 2**8)])
def test_gpu_tn_plan_matches_stepwise_executor_and_oracle(cuda, rows, cols, depth, target):
    import bench
    import torch

    from tensorcircuit_ng_b200 import planner, tnengine

    n = rows * cols
    bits = "".join(str(int(b)) for b in np.random.default_rng(n).integers(0, 2, n))
    td, tensors, sd = _tree(rows, cols, depth, bits, target)
    sliced = list(td["sliced_inds"])
    nsl = int(np.prod([sd[x] for x in sliced])) if sliced else 1
    native = stepwise = 0.0
    with torch.no_grad():
        for s in range(nsl):
            fixed = planner.slice_values(s, sliced, td["size_dict"]) if sliced else None
            tnengine.use_native_plans = True
            a = tnengine.contract_tree(tensors, td["inputs"], td["output"], td["path"], fixed=fixed)
            tnengine.use_native_plans = False
            try:
                b = tnengine.contract_tree(tensors, td["inputs"], td["output"], td["path"], fixed=fixed)
            finally:
                tnengine.use_native_plans = True
            assert abs(complex(a.cpu()) - complex(b.cpu())) <= 1e-7 + 1e-5 * abs(complex(b.cpu()))
            native, stepwise = native + complex(a.cpu()), stepwise + complex(b.cpu())
    ref = bench.build_rcs(otc, rows, cols, depth).wavefunction()[int(bits, 2)]
    assert abs(native - complex(ref)) < 2e-6


@pytest.mark.gpu
def test_gpu_sv_plan_execute_and_vjp_match_per_step_calls(cuda):
    """tcb_sv_plan_execute == the per-pass calls; tcb_sv_plan_vjp == the per-gate adjoint steps."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import _lib, svengine
    from tensorcircuit_ng_b200.passplan import PassStep

    n = 14
    rng = np.random.default_rng(4)
    c = tc.Circuit(n)
    for q in range(n):
        c.h(q)
    for l in range(3):
        for q in range(n - 1):
            c.rzz(q, q + 1, theta=float(rng.uniform(0, 3)))
        for q in range(n):
            c.rx(q, theta=float(rng.uniform(0, 3)))
        c.cnot(0, n - 1)
    nodes, edges = c._copy()
    nn, init, gates = svengine.extract_gate_stream(nodes, edges)
    structure = [(g[1], svengine.gate_kind(g[0], g[2]), int(g[0].tensor.numel())) for g in gates]
    cc = svengine.compile_circuit(nn, structure, torch.device("cuda:0"), absorb_prefix=False)
    gatebuf = svengine.assemble_gatebuf([g[0] for g in gates], torch.device("cuda:0"))
    s1 = svengine.new_zero_state(n, 1, torch.device("cuda:0"))
    cc.run(s1, gatebuf)  # native plan
    s2 = svengine.new_zero_state(n, 1, torch.device("cuda:0"))
    pi = 0
    for step in cc.plan.steps:  # the same passes, one C call each
        assert isinstance(step, PassStep)
        _lib.call("tcb_sv_run_pass", s2.data_ptr(), n, 1, cc.programs.data_ptr() + 4 * cc.offsets[pi], len(step.program),
                  step.tile_bits, step.low_bits, step.pool_elems, gatebuf.data_ptr(), 0, 0, _lib.stream_ptr())  # fmt: skip
        pi += 1
    assert torch.equal(s1, s2)
    want = c.wavefunction()
    assert float((s1 - want.reshape(-1)).abs().max()) < 1e-6
    # vjp: plan walk vs explicit adjoint steps
    from tensorcircuit_ng_b200 import autograd

    tabs = autograd._adjoint_tables(cc, gatebuf)
    src = torch.cat([gatebuf.conj().resolve_conj(), torch.zeros(1, dtype=gatebuf.dtype, device=gatebuf.device)])
    dag = src[tabs.dag_idx].contiguous()
    lam0 = torch.view_as_complex(torch.randn(1 << n, 2, device="cuda"))
    g1 = torch.zeros(tabs.total, 2, dtype=torch.float64, device="cuda")
    lam1, psi1 = lam0.clone(), s1.clone()
    cc.vjp(lam1, psi1, dag, g1)
    g2 = torch.zeros_like(g1)
    lam2, psi2 = lam0.clone(), s1.clone()
    for k, bp, off in reversed(tabs.items):
        _lib.call("tcb_sv_adjoint_step", lam2.data_ptr(), psi2.data_ptr(), n, 1, bp, k, dag.data_ptr() + off * 8, 0,
                  g2.data_ptr() + off * 16, 0, _lib.stream_ptr())  # fmt: skip
    assert torch.equal(lam1, lam2) and torch.equal(psi1, psi2)
    assert float((g1 - g2).abs().max()) <= 1e-9 * float(g2.abs().max())
    zero = torch.zeros(1 << n, dtype=torch.complex64, device="cuda")
    zero[0] = 1
    assert float((psi1 - zero).abs().max()) < 1e-5  # the walk un-computes the state back to |0...0>


@pytest.mark.gpu
def test_gpu_gradient_through_constant_three_qubit_gates(cuda):
    """toffoli / fredkin inside a differentiated circuit: un-applied as constants by the adjoint walk; a
    trainable 3-qubit matrix is refused with a message."""
    import torch

    import tensorcircuit_ng_b200 as tc

    n = 5

    def energy(mod, p, real):
        c = mod.Circuit(n)
        for q in range(n):
            c.h(q)
        for q in range(n):
            c.rx(q, theta=p[q])
        c.toffoli(0, 1, 2)
        for q in range(n - 1):
            c.rzz(q, q + 1, theta=p[n + q])
        c.fredkin(4, 2, 3)
        for q in range(n):
            c.ry(q, theta=p[2 * n - 1 + q])
        return real(c.expectation_ps(z=[0, 3])) + real(c.expectation_ps(x=[2]))

    p0 = np.linspace(0.2, 1.7, 3 * n - 1)
    v, g = tc.backend.value_and_grad(lambda p: energy(tc, p, torch.real))(torch.tensor(p0, dtype=torch.float32))
    f = lambda x: float(energy(otc, x, np.real))
    assert abs(float(v) - f(p0)) < 1e-5
    for k in (0, 3, n + 1, 2 * n + 2):
        xp, xm = p0.copy(), p0.copy()
        xp[k] += 1e-2
        xm[k] -= 1e-2
        assert abs(float(g[k]) - (f(xp) - f(xm)) / 2e-2) < 5e-3
    w = torch.eye(8, dtype=torch.complex64, device="cuda").requires_grad_(True)

    def bad(u):
        c = tc.Circuit(3)
        c.h(0)
        c.any(0, 1, 2, unitary=u)
        return c.expectation_ps(z=[0]).real

    with pytest.raises(tc._lib.EngineError):
        tc.backend.value_and_grad(bad)(w)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [5, 13, 16])
def test_gpu_layered_adjoint_equals_gate_by_gate_walk(cuda, n):
    """Runs of diagonal gates differentiated from one read of the two states (tcb_sv_cross_marginals + fused
    un-apply) give the gradients of the gate-by-gate adjoint walk."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import autograd

    def energy(p):
        c = tc.Circuit(n)
        for q in range(n):
            c.h(q)
        k = 0
        for l in range(2):
            for q in range(n - 1):
                c.rzz(q, q + 1, theta=p[k])
                k += 1
            for q in range(0, n, 2):
                c.rz(q, theta=p[k])
                k += 1
            c.cz(0, n - 1)
            for q in range(n):
                c.rx(q, theta=p[k])
                k += 1
            c.cnot(1, 2)
            c.rz(0, theta=p[k])  # a short diagonal run (walked gate by gate)
            k += 1
        return c.expectation_ps(z=[0, 1]).real + c.expectation_ps(x=[n - 1]).real + c.expectation_ps(y=[2]).real

    nparam = 2 * ((n - 1) + len(range(0, n, 2)) + n + 1)
    p0 = torch.linspace(0.1, 2.9, nparam)
    autograd.layered_adjoint = True
    v1, g1 = tc.backend.value_and_grad(energy)(p0)
    autograd.layered_adjoint = False
    try:
        v2, g2 = tc.backend.value_and_grad(energy)(p0)
    finally:
        autograd.layered_adjoint = True
    assert abs(float(v1) - float(v2)) < 1e-6
    assert float((g1 - g2).abs().max()) < 2e-5
    assert float(g2.abs().max()) > 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("n", [6, 14])
def test_gpu_constant_runs_are_unapplied_fused(cuda, n):
    """A hardware-efficient ansatz (ry / rz layers between CNOT ladders, a toffoli and a swap): the constant gates
    form runs that the backward walk un-applies as fused sub-circuits; gradients equal the gate-by-gate walk and
    central differences of the oracle."""
    import torch

    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import autograd

    def energy(mod, p, real):
        c = mod.Circuit(n)
        k = 0
        for l in range(2):
            for q in range(n):
                c.ry(q, theta=p[k])
                c.rz(q, theta=p[k + 1])
                k += 2
            for q in range(n - 1):
                c.cnot(q, q + 1)
            c.toffoli(0, 2, 4)
            c.swap(1, 3)
            c.h(5)
        return real(c.expectation_ps(z=[0, n - 1])) + real(c.expectation_ps(x=[2])) + real(c.expectation_ps(y=[1], z=[3]))

    p0 = np.linspace(0.15, 2.6, 4 * n)
    pt = torch.tensor(p0, dtype=torch.float32)
    seen = []
    orig = autograd._adjoint_tables

    def spy(cc, gb, mask=None):
        t = orig(cc, gb, mask)
        seen.append(t)
        return t

    autograd._adjoint_tables = spy
    try:
        v1, g1 = tc.backend.value_and_grad(lambda x: energy(tc, x, torch.real))(pt)
    finally:
        autograd._adjoint_tables = orig
    assert any(isinstance(sg, autograd._ConstRun) for t in seen for sg in t.segments)
    # the ry(q) rz(q) blocks repeat qubits: split into rounds of one-qubit gates on distinct qubits
    assert any(isinstance(sg, autograd._OneQubitRun) and sg.first == -1 for t in seen for sg in t.segments)
    autograd.layered_adjoint = False
    try:
        v2, g2 = tc.backend.value_and_grad(lambda x: energy(tc, x, torch.real))(pt)
    finally:
        autograd.layered_adjoint = True
    assert abs(float(v1) - float(v2)) < 1e-6 and float((g1 - g2).abs().max()) < 2e-5
    f = lambda x: float(energy(otc, x, np.real))
    assert abs(float(v1) - f(p0)) < 1e-5
    for k in (0, 3, 2 * n + 1, 4 * n - 1):
        xp, xm = p0.copy(), p0.copy()
        xp[k] += 1e-2
        xm[k] -= 1e-2
        assert abs(float(g1[k]) - (f(xp) - f(xm)) / 2e-2) < 5e-3
