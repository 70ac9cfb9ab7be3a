"""
ORACLE (test infrastructure, never imported by the product package).

numpy restatement of the gate tensors of TensorCircuit-NG
(/root/reference/tensorcircuit/gates.py).  Every function cites the lines it
follows.  Layout: a k-qubit gate is a rank-2k tensor [out_0..out_{k-1},
in_0..in_{k-1}] (gates.py:497-516 reshapes 4x4/8x8 matrices row-major).
All tensors are complex64 (cons.py:74, dtypestr).
"""

from __future__ import annotations

from functools import reduce
from operator import mul
from typing import Any, Optional, Sequence, Union

import numpy as np
import scipy.linalg

from . import tn

npdtype = np.complex64


class Gate(tn.Node):  # gates.py:185-224
    def copy(self, conjugate: bool = False) -> "Gate":
        r = super().copy(conjugate)
        r.__class__ = Gate
        return r


# gates.py:33-174 -------------------------------------------------------------
_i00 = np.array([[1.0, 0.0], [0.0, 0.0]])
_i01 = np.array([[0.0, 1.0], [0.0, 0.0]])
_i10 = np.array([[0.0, 0.0], [1.0, 0.0]])
_i11 = np.array([[0.0, 0.0], [0.0, 1.0]])

_h_matrix = 1 / np.sqrt(2) * np.array([[1.0, 1.0], [1.0, -1.0]])
_i_matrix = np.array([[1.0, 0.0], [0.0, 1.0]])
_x_matrix = np.array([[0.0, 1.0], [1.0, 0.0]])
_y_matrix = np.array([[0.0, -1j], [1j, 0.0]])
_z_matrix = np.array([[1.0, 0.0], [0.0, -1.0]])
_s_matrix = np.array([[1.0, 0.0], [0.0, 1j]])
_t_matrix = np.array([[1.0, 0.0], [0.0, np.exp(np.pi / 4 * 1j)]])
_wroot_matrix = (
    1 / np.sqrt(2) * np.array([[1, -1 / np.sqrt(2) * (1 + 1.0j)], [1 / np.sqrt(2) * (1 - 1.0j), 1]])
)
_xx_matrix = np.kron(_x_matrix, _x_matrix)
_yy_matrix = np.kron(_y_matrix, _y_matrix)
_zz_matrix = np.kron(_z_matrix, _z_matrix)

_cnot_matrix = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0], [0, 0, 1.0, 0]])
_cz_matrix = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, -1.0]])
_cy_matrix = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, -1.0j], [0, 0, 1.0j, 0]])
_swap_matrix = np.array([[1.0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]])
_iswap_d1_matrix = np.diag([1.0, 0, 0, 1.0])
_iswap_d2_matrix = np.diag([0, 1.0, 1.0, 0])
_iswap_od_matrix = np.array([[0, 0, 0, 0], [0, 0, 1.0, 0], [0, 1.0, 0, 0], [0, 0, 0, 0]])
_toffoli_matrix = np.eye(8)[[0, 1, 2, 3, 4, 5, 7, 6]]
_fredkin_matrix = np.eye(8)[[0, 1, 2, 3, 4, 6, 5, 7]]


def _t(m: Any) -> np.ndarray:
    return np.asarray(m).astype(npdtype)


def _reshape2(m: np.ndarray) -> np.ndarray:
    n = int(round(np.log2(m.size)))
    return np.reshape(m, [2] * n)


def _fixed(m: np.ndarray, name: str):  # gates.py:497-516 meta_gate / GateF.__call__ :298-312
    def f() -> Gate:
        return Gate(_reshape2(_t(m)), name=name)

    f.n = name  # type: ignore[attr-defined]
    return f


i = _fixed(_i_matrix, "i")
x = _fixed(_x_matrix, "x")
y = _fixed(_y_matrix, "y")
z = _fixed(_z_matrix, "z")
h = _fixed(_h_matrix, "h")
s = _fixed(_s_matrix, "s")
t = _fixed(_t_matrix, "t")
sd = _fixed(np.conj(_s_matrix).T, "sd")  # gates.py:1217-1219 adjoint()
td = _fixed(np.conj(_t_matrix).T, "td")
wroot = _fixed(_wroot_matrix, "wroot")
cnot = _fixed(_cnot_matrix, "cnot")
cz = _fixed(_cz_matrix, "cz")
cy = _fixed(_cy_matrix, "cy")
swap = _fixed(_swap_matrix, "swap")
toffoli = _fixed(_toffoli_matrix, "toffoli")
fredkin = _fixed(_fredkin_matrix, "fredkin")


def _controlled(u: np.ndarray) -> np.ndarray:  # gates.py:343-360
    u = np.reshape(u, [int(np.sqrt(u.size))] * 2)
    s_ = u.shape[-1]
    cu = np.block([[np.eye(s_), np.zeros([s_, s_])], [np.zeros([s_, s_]), u]])
    return _reshape2(_t(cu))


def _ocontrolled(u: np.ndarray) -> np.ndarray:  # gates.py:362-380
    u = np.reshape(u, [int(np.sqrt(u.size))] * 2)
    s_ = u.shape[-1]
    cu = np.block([[u, np.zeros([s_, s_])], [np.zeros([s_, s_]), np.eye(s_)]])
    return _reshape2(_t(cu))


ox = lambda: Gate(_ocontrolled(_x_matrix), name="ox")
oy = lambda: Gate(_ocontrolled(_y_matrix), name="oy")
oz = lambda: Gate(_ocontrolled(_z_matrix), name="oz")


def phase_gate(theta: float = 0) -> Gate:  # gates.py:584-603
    theta = _t(theta)
    return Gate(_t(_i00) + np.exp(1.0j * theta) * _t(_i11))


def u_gate(theta: float = 0.0, phi: float = 0.0, lbd: float = 0.0) -> Gate:  # gates.py:630-658
    theta, phi, lbd = _t(theta), _t(phi), _t(lbd)
    unitary = (
        np.cos(theta / 2) * _t(_i00)
        - np.exp(1.0j * lbd) * np.sin(theta / 2) * _t(_i01)
        + np.exp(1.0j * phi) * np.sin(theta / 2) * _t(_i10)
        + np.exp(1.0j * (phi + lbd)) * np.cos(theta / 2) * _t(_i11)
    )
    return Gate(_t(unitary))


def r_gate(theta: float = 0.0, alpha: float = 0.0, phi: float = 0.0) -> Gate:  # gates.py:661-689
    theta, phi, alpha = _t(theta), _t(phi), _t(alpha)
    unitary = (
        np.cos(theta) * _t(_i_matrix)
        - 1.0j * np.cos(phi) * np.sin(alpha) * np.sin(theta) * _t(_x_matrix)
        - 1.0j * np.sin(phi) * np.sin(alpha) * np.sin(theta) * _t(_y_matrix)
        - 1.0j * np.sin(theta) * np.cos(alpha) * _t(_z_matrix)
    )
    return Gate(_t(unitary))


def rx_gate(theta: float = 0.0) -> Gate:  # gates.py:692-707
    theta = _t(theta)
    return Gate(_t(np.cos(theta / 2.0) * _t(_i_matrix) - 1.0j * np.sin(theta / 2.0) * _t(_x_matrix)))


def ry_gate(theta: float = 0.0) -> Gate:  # gates.py:710-725
    theta = _t(theta)
    return Gate(_t(np.cos(theta / 2.0) * _t(_i_matrix) - 1.0j * np.sin(theta / 2.0) * _t(_y_matrix)))


def rz_gate(theta: float = 0.0) -> Gate:  # gates.py:728-743
    theta = _t(theta)
    return Gate(_t(np.cos(theta / 2.0) * _t(_i_matrix) - 1.0j * np.sin(theta / 2.0) * _t(_z_matrix)))


def iswap_gate(theta: float = 1.0) -> Gate:  # gates.py:788-814
    theta = _t(theta)
    unitary = (
        _t(_iswap_d1_matrix)
        + np.cos(theta * np.pi / 2) * _t(_iswap_d2_matrix)
        + 1.0j * np.sin(theta * np.pi / 2) * _t(_iswap_od_matrix)
    )
    return Gate(np.reshape(_t(unitary), [2, 2, 2, 2]))


def cr_gate(theta: float = 0.0, alpha: float = 0.0, phi: float = 0.0) -> Gate:  # gates.py:817-849
    theta, phi, alpha = _t(theta), _t(phi), _t(alpha)
    u = np.array([[1.0, 0.0], [0.0, 0.0]])
    d = np.array([[0.0, 0.0], [0.0, 1.0]])
    j = _t(np.kron(u, _i_matrix))
    i_ = _t(np.kron(d, _i_matrix))
    x_ = _t(np.kron(d, _x_matrix))
    y_ = _t(np.kron(d, _y_matrix))
    z_ = _t(np.kron(d, _z_matrix))
    unitary = (
        j
        + np.cos(theta) * i_
        - 1.0j * np.cos(phi) * np.sin(alpha) * np.sin(theta) * x_
        - 1.0j * np.sin(phi) * np.sin(alpha) * np.sin(theta) * y_
        - 1.0j * np.sin(theta) * np.cos(alpha) * z_
    )
    return Gate(np.reshape(_t(unitary), [2, 2, 2, 2]))


def any_gate(unitary: Any, name: str = "any") -> Gate:  # gates.py:866-890
    if isinstance(unitary, tn.Node):
        unitary.tensor = _t(unitary.tensor)
        if not isinstance(unitary, Gate):
            unitary.__class__ = Gate
        return unitary  # type: ignore[return-value]
    return Gate(_reshape2(_t(unitary)), name=name)


def exponential_gate(unitary: Any, theta: float, name: str = "none") -> Gate:  # gates.py:893-914
    theta, unitary = _t(theta), _t(unitary)
    d = int(np.sqrt(unitary.size))
    mat = scipy.linalg.expm(-1.0j * theta * np.reshape(unitary, [d, d]).astype(np.complex128))
    return Gate(_reshape2(_t(mat)), name="exp-" + name)


def exponential_gate_unity(unitary: Any, theta: float, half: bool = False, name: str = "none") -> Gate:
    # gates.py:920-953: cos(theta) I - i sin(theta) U, valid for U^2 = I
    theta, unitary = _t(theta), _t(unitary)
    n = int(np.log2(unitary.size))
    it = _t(np.eye(2 ** (n // 2)).reshape([2] * n))
    unitary = np.reshape(unitary, [2] * n)
    if half is True:
        theta = theta / 2.0
    mat = np.cos(theta) * it - 1.0j * np.sin(theta) * unitary
    return Gate(_t(mat), name="exp1-" + name)


exp_gate = exponential_gate
exp1_gate = exponential_gate_unity


def rzz_gate(theta: float = 0.0) -> Gate:  # gates.py:976
    return exp1_gate(_zz_matrix, theta, half=True)


def rxx_gate(theta: float = 0.0) -> Gate:  # gates.py:977
    return exp1_gate(_xx_matrix, theta, half=True)


def ryy_gate(theta: float = 0.0) -> Gate:  # gates.py:978
    return exp1_gate(_yy_matrix, theta, half=True)


def _ctl(f, name):  # gates.py:1212-1214  getattr(thismodule, f[1:]).controlled()
    def g(**kws: Any) -> Gate:
        return Gate(_controlled(f(**kws).tensor), name=name)

    return g


def _octl(f, name):  # gates.py:1215-1217
    def g(**kws: Any) -> Gate:
        return Gate(_ocontrolled(f(**kws).tensor), name=name)

    return g


cu_gate = _ctl(u_gate, "cu")
crx_gate = _ctl(rx_gate, "crx")
cry_gate = _ctl(ry_gate, "cry")
crz_gate = _ctl(rz_gate, "crz")
cphase_gate = _ctl(phase_gate, "cphase")
orx_gate = _octl(rx_gate, "orx")
ory_gate = _octl(ry_gate, "ory")
orz_gate = _octl(rz_gate, "orz")


def diagonal_gate(diag: Any, dim: int = 2, name: str = "diagonal") -> Gate:  # gates.py:1059-1078
    diag = _t(diag)
    noe = int(np.round(np.log(diag.size) / np.log(dim)))
    return Gate(np.reshape(diag, [dim] * noe), name=name)


def matrix_for_gate(g: Gate) -> np.ndarray:
    t_ = g.tensor
    d = int(np.sqrt(t_.size))
    return np.reshape(t_, [d, d])
