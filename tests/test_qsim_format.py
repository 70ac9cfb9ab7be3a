"""SURVEY §8f rank 3 — Google qsim circuit files (`Circuit.from_qsim_file`, abstractcircuit.py:1269-1351), the
public RCS input format that feeds config 5.  The file of the reference's tests/test_circuit.py:2443-2500."""
import numpy as np
import pytest

import tc_oracle as otc
from tc_oracle.circuit import apply_qsim

LINES = ["2", "0 h 0", "1 cnot 0 1", "2 x 1", "3 y 0", "4 z 1", "5 s 0", "6 t 1", "7 x_1_2 0", "8 y_1_2 1",
         "9 z_1_2 0", "10 w_1_2 1", "11 hz_1_2 0", "12 cz 0 1", "13 is 0 1", "14 rx 0 0.7", "15 ry 1 0.3",
         "16 rz 0 0.5", "17 fsim 0 1 0.4 0.2"]  # fmt: skip
NAMES = ["h", "cnot", "x", "y", "z", "phase", "u", "wroot", "iswap", "cphase", "rx", "ry", "rz", "cz"]


def _rcs_lines(rows, cols, depth, seed):
    """A small Sycamore-style file: random sqrt gates + fsim couplers in the qsim vocabulary."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    out = [str(n)]
    for d in range(depth):
        for q in range(n):
            out.append(f"{2 * d} {['x_1_2', 'y_1_2', 'hz_1_2', 'w_1_2'][int(rng.integers(0, 4))]} {q}")
        for r in range(rows):
            for c in range(cols - 1):
                if (c + d) % 2 == 0:
                    a, b = r * cols + c, r * cols + c + 1
                    out.append(f"{2 * d + 1} fs {a} {b} {rng.uniform(0.2, 1.4):.6f} {rng.uniform(0.1, 0.6):.6f}"
                               if d % 2 else f"{2 * d + 1} cz {a} {b}")  # fmt: skip
    return out


def test_oracle_qsim_roundtrip():
    c = apply_qsim(otc.Circuit(2), LINES)
    assert len(c._qir) == (len(LINES) - 1) + 1  # fsim -> iswap + cphase
    m = c.matrix()
    np.testing.assert_allclose(m @ m.conj().T, np.eye(4), atol=1e-5)
    with pytest.raises(NotImplementedError):
        apply_qsim(otc.Circuit(2), ["2", "0 bogus 0"])


def test_engine_parses_qsim_like_the_reference(tmp_path):
    import tensorcircuit_ng_b200 as tc

    path = tmp_path / "c.qsim"
    path.write_text("\n".join(LINES))
    c = tc.Circuit.from_qsim_file(str(path))
    assert c._nqubits == 2 and len(c._qir) == (len(LINES) - 1) + 1
    names = [d["name"] for d in c._qir]
    for g in NAMES:
        assert g in names
    onames = [d["name"] for d in apply_qsim(otc.Circuit(2), LINES)._qir]
    assert names == onames
    with pytest.raises(NotImplementedError):
        tc.Circuit._apply_qsim(tc.Circuit(2), ["2", "0 bogus 0"])


@pytest.mark.gpu
def test_gpu_qsim_circuits_match_oracle(cuda, tmp_path):
    import tensorcircuit_ng_b200 as tc

    path = tmp_path / "c.qsim"
    path.write_text("\n".join(LINES))
    c = tc.Circuit.from_qsim_file(str(path))
    want = apply_qsim(otc.Circuit(2), LINES)
    assert np.abs(c.wavefunction().cpu().numpy() - want.wavefunction()).max() < 1e-6
    m = c.matrix().cpu().numpy()
    np.testing.assert_allclose(m @ m.conj().T, np.eye(4), atol=1e-5)
    assert np.abs(m - want.matrix()).max() < 1e-5
    for rows, cols, depth, seed in [(3, 3, 6, 0), (3, 4, 8, 1), (4, 4, 5, 2)]:
        lines = _rcs_lines(rows, cols, depth, seed)
        p = tmp_path / f"rcs{seed}.qsim"
        p.write_text("\n".join(lines))
        got = tc.Circuit.from_qsim_file(str(p)).wavefunction().cpu().numpy()
        ref = apply_qsim(otc.Circuit(rows * cols), lines).wavefunction()
        assert np.abs(got - ref).max() < 2e-6
        bits = "".join(str(int(b)) for b in np.random.default_rng(seed).integers(0, 2, rows * cols))
        amp = complex(tc.Circuit.from_qsim_file(str(p)).amplitude(bits).cpu())
        assert abs(amp - ref[int(bits, 2)]) < 2e-6
