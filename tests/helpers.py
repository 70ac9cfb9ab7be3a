"""Shared test helpers: identical circuits built on the oracle (numpy) and on the product (torch)."""
import numpy as np


def random_layers(n, depth, seed):
    """A gate list [(name, qubits, kwargs)] exercising every kernel path."""
    rng = np.random.default_rng(seed)
    ops = []
    for d in range(depth):
        for q in range(n):
            r = int(rng.integers(0, 7))
            th = float(rng.uniform(0, 2 * np.pi))
            if r == 0:
                ops.append(("rx", [q], {"theta": th}))
            elif r == 1:
                ops.append(("ry", [q], {"theta": th}))
            elif r == 2:
                ops.append(("rz", [q], {"theta": th}))
            elif r == 3:
                ops.append(("h", [q], {}))
            elif r == 4:
                ops.append(("phase", [q], {"theta": th}))
            elif r == 5:
                ops.append(("t", [q], {}))
            else:
                ops.append(("u", [q], {"theta": th, "phi": th / 3, "lbd": -th / 5}))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            a, b = int(perm[i]), int(perm[i + 1])
            r = int(rng.integers(0, 9))
            th = float(rng.uniform(0, 2 * np.pi))
            if r == 0:
                ops.append(("cz", [a, b], {}))
            elif r == 1:
                ops.append(("cnot", [a, b], {}))
            elif r == 2:
                ops.append(("rzz", [a, b], {"theta": th}))
            elif r == 3:
                ops.append(("crx", [a, b], {"theta": th}))
            elif r == 4:
                ops.append(("rxx", [a, b], {"theta": th}))
            elif r == 5:
                ops.append(("ox", [a, b], {}))
            elif r == 6:
                ops.append(("iswap", [a, b], {"theta": th}))
            elif r == 7:
                ops.append(("cphase", [a, b], {"theta": th}))
            else:
                ops.append(("swap", [a, b], {}))
        if n >= 3 and d % 2 == 0:
            a, b, c = [int(v) for v in rng.permutation(n)[:3]]
            ops.append(("toffoli", [a, b, c], {}))
            ops.append(("fredkin", [c, a, b], {}))
    return ops


def build(mod, n, ops, inputs=None):
    c = mod.Circuit(n) if inputs is None else mod.Circuit(n, inputs=inputs)
    for name, qs, kw in ops:
        getattr(c, name)(*qs, **kw)
    return c


def oracle_circuit(n, ops, inputs=None):
    import tc_oracle

    return build(tc_oracle, n, ops, inputs)


def oracle_state(n, ops, inputs=None, contractor=None):
    """Oracle wavefunction.  The reference's default greedy contractor is used for small
    registers; wider ones use the reference's literal statevector order
    (`plain_contractor`, tensorcircuit/cons.py:429-463) because greedy paths on non-1D
    circuits (e.g. QAOA on a 3-regular graph) build intermediates far larger than 2^n."""
    from tc_oracle import cons

    method = contractor or ("greedy" if n <= 12 else "plain")
    kws = {"preprocessing": True} if method == "greedy" else {}
    with cons.runtime_contractor(method, **kws):
        return oracle_circuit(n, ops, inputs).wavefunction()


def brickwork(n, depth, seed=0):
    """BASELINE.json config 1: rx on every qubit, rzz on alternating pairs."""
    rng = np.random.default_rng(seed)
    ops = []
    for l in range(depth):
        for q in range(n):
            ops.append(("rx", [q], {"theta": float(rng.uniform(0, 2 * np.pi))}))
        for q in range(l % 2, n - 1, 2):
            ops.append(("rzz", [q, q + 1], {"theta": float(rng.uniform(0, 2 * np.pi))}))
    return ops


def qaoa(n, p, seed=0):
    """BASELINE.json config 3: QAOA MaxCut on a random 3-regular graph."""
    import networkx as nx

    g = nx.random_regular_graph(3, n, seed=seed)
    rng = np.random.default_rng(seed)
    gammas = rng.uniform(0, np.pi, size=p)
    betas = rng.uniform(0, np.pi, size=p)
    ops = [("h", [q], {}) for q in range(n)]
    from tc_oracle import gates as og

    for l in range(p):
        for a, b in g.edges:
            ops.append(("exp1", [int(a), int(b)], {"unitary": og._zz_matrix, "theta": float(gammas[l])}))
        for q in g.nodes:
            ops.append(("rx", [int(q)], {"theta": float(betas[l])}))
    return ops, [(int(a), int(b)) for a, b in g.edges]


def param_shift(f, params, idx, kind="half"):
    """Exact derivative of an expectation value with respect to ONE gate parameter by the parameter-shift rule,
    evaluated on the oracle: gates exp(-i theta/2 P) (rx, ry, rz, rzz, rxx, ryy; `kind="half"`) shift by pi/2 and
    halve, gates exp(-i theta P) (exp1 with P^2 = 1; `kind="full"`) shift by pi/4.  The parameter must feed exactly
    one gate.  No step-size error: the oracle's complex64 rounding (~1e-6) is all that is left."""
    s = np.pi / 2 if kind == "half" else np.pi / 4
    pp, pm = np.array(params, dtype=np.float64), np.array(params, dtype=np.float64)
    pp[idx] += s
    pm[idx] -= s
    d = f(pp) - f(pm)
    return d / 2 if kind == "half" else d
