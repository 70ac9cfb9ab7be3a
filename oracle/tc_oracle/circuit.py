"""
ORACLE (test infrastructure, never imported by the product package).

numpy restatement of the Circuit front end of TensorCircuit-NG on the hot path:
  /root/reference/tensorcircuit/circuit.py:44-131 (init), :701-721 (wavefunction),
      :833-913 (expectation)
  /root/reference/tensorcircuit/basecircuit.py:52-66, :151-181, :183-371,
      :375-447, :562-624
  /root/reference/tensorcircuit/abstractcircuit.py:37-81, :114-240, :1523-1603
  /root/reference/tensorcircuit/simplify.py:198-296 (light cone)
  /root/reference/tensorcircuit/quantum.py:166-183 (onehot), :1475-1495 (ps2xyz)
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import cons, gates, tn
from .cons import contractor  # rebound by cons._set_global_contractor
from .gates import Gate

npdtype = np.complex64


def ps2xyz(ps: Sequence[int]) -> Dict[str, List[int]]:  # quantum.py:1475-1495
    xyz: Dict[str, List[int]] = {"x": [], "y": [], "z": []}
    for i, j in enumerate(ps):
        if j == 1:
            xyz["x"].append(i)
        if j == 2:
            xyz["y"].append(i)
        if j == 3:
            xyz["z"].append(i)
    return xyz


def _is_seq(x: Any) -> bool:
    return isinstance(x, (list, tuple, range, np.ndarray))


# simplify.py:198-296 -------------------------------------------------------------
def _light_cone_cancel(nodes: List[Any]) -> Tuple[List[Any], bool]:
    is_changed = False
    nodes_to_remove = set()
    for i in range(len(nodes) - 1, -1, -1):
        n = nodes[i]
        if n in nodes_to_remove:
            continue
        if getattr(n, "is_dagger", None) is True:
            continue
        noe = len(n.shape)
        if noe % 2 != 0:
            continue
        match_node = None
        for leg_idx in range(noe // 2):
            e = n[leg_idx]
            if e.is_dangling():
                break
            n1, n2 = e.node1, e.node2
            other = n2 if n1 is n else n1
            if getattr(other, "is_dagger", None) is not True:
                break
            if getattr(other, "id", None) != getattr(n, "id", -1):
                break
            if e.axis1 != e.axis2:
                break
            if match_node is None:
                match_node = other
            elif match_node is not other:
                break
        else:
            if match_node is not None and match_node not in nodes_to_remove:
                for leg_idx in range(noe // 2, noe):
                    e_n = n[leg_idx]
                    e_m = match_node[leg_idx]
                    m_n, i_n = (e_n.node2, e_n.axis2) if e_n.node1 is n else (e_n.node1, e_n.axis1)
                    m_m, i_m = (
                        (e_m.node2, e_m.axis2) if e_m.node1 is match_node else (e_m.node1, e_m.axis1)
                    )
                    e_n.disconnect()
                    e_m.disconnect()
                    m_n[i_n] ^ m_m[i_m]
                nodes_to_remove.add(n)
                nodes_to_remove.add(match_node)
                is_changed = True
    if is_changed:
        return [n for n in nodes if n not in nodes_to_remove], True
    return nodes, False


def _full_light_cone_cancel(nodes: List[Any]) -> List[Any]:
    if not nodes:
        return nodes
    if any(getattr(n, "is_dagger", None) is None for n in nodes):
        return nodes
    nodes, is_changed = _light_cone_cancel(nodes)
    while is_changed:
        nodes, is_changed = _light_cone_cancel(nodes)
    return nodes


class Circuit:
    # abstractcircuit.py:37-81
    sgates = (
        ["i", "x", "y", "z", "h", "t", "s", "td", "sd", "wroot"]
        + ["cnot", "cz", "swap", "cy", "ox", "oy", "oz"]
        + ["toffoli", "fredkin"]
    )
    vgates = [
        "r", "cr", "u", "cu", "rx", "ry", "rz", "phase", "rxx", "ryy", "rzz", "cphase",
        "crx", "cry", "crz", "orx", "ory", "orz", "iswap", "any", "exp", "exp1",
    ]  # fmt: skip
    gate_aliases = [["cnot", "cx"], ["fredkin", "cswap"], ["toffoli", "ccnot"], ["toffoli", "ccx"],
                    ["any", "unitary"], ["sd", "sdg"], ["td", "tdg"]]  # fmt: skip
    is_dm = False

    def __init__(self, nqubits: int, inputs: Optional[Any] = None) -> None:  # circuit.py:44-131
        self._nqubits = nqubits
        self._d = 2
        if inputs is None:
            nodes = self.all_zero_nodes(nqubits)
            self._front = [n.get_edge(0) for n in nodes]
        else:
            inputs = np.asarray(inputs).astype(npdtype).reshape([-1])
            N = inputs.shape[0]
            n = int(round(np.log2(N)))
            if n != nqubits and n != 2 * nqubits:
                raise ValueError(
                    f"inputs has {N} elements => {n} sites (dim=2), "
                    f"expected {nqubits} (state) or {2 * nqubits} (density matrix)"
                )
            inp = Gate(np.reshape(inputs, [2] * n))
            nodes = [inp]
            self._front = [inp.get_edge(i) for i in range(n)]
        self.coloring_nodes(nodes, flag="inputs")
        self._nodes = nodes
        self._start_index = len(nodes)
        self._qir: List[Dict[str, Any]] = []
        self.state_tensor = None

    @staticmethod
    def all_zero_nodes(n: int, prefix: str = "qb-") -> List[tn.Node]:  # basecircuit.py:52-66
        return [tn.Node(np.array([1.0, 0.0], dtype=npdtype), name=prefix + str(x)) for x in range(n)]

    @staticmethod
    def coloring_nodes(nodes, is_dagger: bool = False, flag: str = "inputs") -> None:  # :105-123
        for node in nodes:
            node.is_dagger = is_dagger
            node.flag = flag
            node.id = id(node)

    @staticmethod
    def copy_nodes(nodes, dangling=None, conj: Optional[bool] = False):  # basecircuit.py:151-176
        ndict, edict = tn.copy(nodes, conjugate=conj)
        newnodes = []
        for n in nodes:
            newn = ndict[n]
            newn.is_dagger = conj
            newn.flag = getattr(n, "flag", "") + "copy"
            newn.id = getattr(n, "id", id(n))
            newnodes.append(newn)
        newfront = []
        if not dangling:
            dangling = []
            for n in nodes:
                dangling.extend([e for e in n])
        for e in dangling:
            newfront.append(edict[e])
        return newnodes, newfront

    def _copy(self, conj: Optional[bool] = False):  # basecircuit.py:178-181
        return self.copy_nodes(self._nodes, self._front, conj)

    # basecircuit.py:183-371 (dense and diagonal branches; split/mpo are out of scope)
    def apply_general_gate(
        self, gate: Gate, *index: int, name: Optional[str] = None, diagonal: bool = False, **_: Any
    ) -> None:
        if name is None:
            name = ""
        self._qir.append({"gate": gate, "index": index, "name": name, "diagonal": diagonal})
        if len(index) != len(set(index)):
            raise ValueError(
                f"gate index {list(index)} has duplicate qubits; " "each qubit may appear at most once"
            )
        index = tuple(i if i >= 0 else self._nqubits + i for i in index)
        noe = len(index)
        if not diagonal:
            gate.name = name
            self.coloring_nodes([gate], flag="gate")
            self._nodes.append(gate)
            for i, ind in enumerate(index):
                gate.get_edge(i + noe) ^ self._front[ind]
                self._front[ind] = gate.get_edge(i)
        else:
            self.coloring_nodes([gate], flag="gate")
            gate.id = id(gate)
            gate.name = name
            self._nodes.append(gate)
            for i, ind in enumerate(index):
                phys_edge = gate[i]
                cn = tn.CopyNode(3, self._d, name=f"{name}_copy_{i}")
                self.coloring_nodes([cn], flag="gate")
                self._nodes.append(cn)
                cn[0] ^ self._front[ind]
                cn[1] ^ phys_edge
                self._front[ind] = cn[2]
        self.state_tensor = None

    apply = apply_general_gate

    def _apply_named(self, gname: str, *index: Any, **vars: Any) -> None:
        # abstractcircuit.py:114-240 apply / apply_list index broadcasting
        if isinstance(index[0], (int, np.integer)):
            self._apply_one(gname, *index, **vars)
        elif _is_seq(index[0]):
            for i, ind in enumerate(zip(*index)):
                nvars = {}
                for k, v in vars.items():
                    try:
                        nvars[k] = v[i]
                    except Exception:  # pylint: disable=broad-except
                        nvars[k] = v
                self._apply_one(gname, *ind, **nvars)
        else:
            raise ValueError("Illegal index specification")

    def _apply_one(self, gname: str, *index: int, **vars: Any) -> None:
        localname = vars.pop("name", gname)
        vars.pop("split", None)
        if gname in self.sgates:
            gate = getattr(gates, gname)()
        elif gname == "diagonal":
            gate = gates.diagonal_gate(**vars)
            self.apply_general_gate(gate, *index, name=localname, diagonal=True)
            return
        else:
            f = getattr(gates, gname + "_gate")
            gate = f(**vars)
        self.apply_general_gate(gate, *index, name=localname)

    # circuit.py:701-721
    def wavefunction(self, form: str = "default") -> np.ndarray:
        nodes, d_edges = self._copy()
        t = contractor(nodes, output_edge_order=d_edges)
        shape = {"default": [-1], "ket": [-1, 1], "bra": [1, -1]}[form]
        return np.reshape(t.tensor, shape)

    state = wavefunction

    def matrix(self) -> np.ndarray:  # circuit.py:743-769 (identity inputs)
        n = self._nqubits
        c = Circuit(n, inputs=np.eye(2**n))
        # replay: front [0..n) are outputs, [n..2n) inputs of the identity
        for d in self._qir:
            c.apply_general_gate(
                Gate(d["gate"].tensor), *d["index"], name=d["name"], diagonal=d["diagonal"]
            )
        return np.reshape(c.state(), [2**n, 2**n])

    # basecircuit.py:375-391
    def _copy_state_tensor(self, conj: bool = False, reuse: bool = True):
        if reuse:
            t = getattr(self, "state_tensor", None)
            if t is None:
                nodes, d_edges = self._copy()
                t = contractor(nodes, output_edge_order=d_edges)
                setattr(self, "state_tensor", t)
            ndict, edict = tn.copy([t], conjugate=conj)
            newnodes = [ndict[t]]
            newfront = [edict[e] for e in t.edges]
            return newnodes, newfront
        return self._copy(conj)

    # basecircuit.py:393-447
    def expectation_before(self, *ops: Tuple[Any, Any], reuse: bool = True, **kws: Any) -> List[tn.Node]:
        nq = self._nqubits
        nodes1, edge1 = self._copy_state_tensor(reuse=reuse)
        nodes2, edge2 = self._copy_state_tensor(conj=True, reuse=reuse)
        nodes = nodes1 + nodes2
        newdang = edge1 + edge2
        occupied = set()
        for op, index in ops:
            if not isinstance(op, tn.Node):
                op = np.asarray(op).astype(npdtype)
                op = Gate(np.reshape(op, [2] * int(round(np.log2(op.size)))))
            else:
                op.tensor = op.tensor.astype(npdtype)
            if isinstance(index, (int, np.integer)):
                index = [index]
            index = tuple(i if i >= 0 else self._nqubits + i for i in index)
            noe = len(index)
            for j, e in enumerate(index):
                if e in occupied:
                    raise ValueError(
                        f"Cannot measure two operators in one index: qubit {e} "
                        f"is already occupied by a previous operator in this "
                        f"measurement, index={index}"
                    )
                newdang[e + nq] ^ op.get_edge(j)
                newdang[e] ^ op.get_edge(j + noe)
                occupied.add(e)
            self.coloring_nodes([op], flag="operator")
            nodes.append(op)
        for j in range(nq):
            if j not in occupied:
                newdang[j] ^ newdang[j + nq]
        return nodes

    # circuit.py:833-913
    def expectation(self, *ops: Tuple[Any, Any], reuse: bool = True, enable_lightcone: bool = False, **kws):
        if enable_lightcone:
            reuse = False
        nodes1 = self.expectation_before(*ops, reuse=reuse)
        if enable_lightcone:
            nodes1 = _full_light_cone_cancel(nodes1)
        return contractor(nodes1).tensor

    # abstractcircuit.py:1523-1603
    def expectation_ps(self, x=None, y=None, z=None, ps=None, reuse: bool = True, **kws: Any):
        obs = []
        if ps is not None:
            d = ps2xyz(ps)
            x, y, z = d.get("x", None), d.get("y", None), d.get("z", None)
        if x is not None:
            for i in x:
                obs.append([gates.x(), [i]])
        if y is not None:
            for i in y:
                obs.append([gates.y(), [i]])
        if z is not None:
            for i in z:
                obs.append([gates.z(), [i]])
        return self.expectation(*obs, reuse=reuse, **kws)

    # basecircuit.py:562-624
    def amplitude_before(self, l: Union[str, Sequence[int]]) -> List[tn.Node]:
        no, d_edges = self._copy()
        if isinstance(l, str):
            l = [int(ch) for ch in l]
        l = np.asarray(l).astype(np.int32)
        endns = np.eye(2, dtype=npdtype)[l]  # quantum.py:166-183 onehot_d_tensor
        ms = []
        for i in range(self._nqubits):
            n = tn.Node(endns[i])
            self.coloring_nodes([n], flag="measurement")
            ms.append(n)
            d_edges[i] ^ n.get_edge(0)
        no.extend(ms)
        return no

    def amplitude(self, l: Union[str, Sequence[int]]):
        no = self.amplitude_before(l)
        return contractor(no).tensor


def apply_qsim(c: "Circuit", lines: Sequence[str]) -> "Circuit":
    """abstractcircuit.py:1284-1351: Google qsim text -> gates (line 0 = qubit count, then
    "<moment> <gate> <qubits> [<params>]"); the alias table of `:1306-1345`."""
    half, quarter = np.pi / 2, np.pi / 4
    for ln in lines[1:]:
        t = ln.strip().split(" ")
        if len(t) < 2:
            continue
        g, a = t[1].lower(), t[2:]
        if g in ("h", "x", "y", "z"):
            getattr(c, g)(int(a[0]))
        elif g in ("s", "t"):
            c.phase(int(a[0]), theta=half if g == "s" else quarter)
        elif g in ("x_1_2", "y_1_2", "z_1_2"):
            getattr(c, "r" + g[0])(int(a[0]), theta=half)
        elif g == "w_1_2":
            c.u(int(a[0]), theta=half, phi=-quarter, lbd=quarter)
        elif g == "hz_1_2":
            c.wroot(int(a[0]))
        elif g in ("cnot", "cx", "cy", "cz"):
            getattr(c, "cnot" if g == "cx" else g)(int(a[0]), int(a[1]))
        elif g in ("is", "iswap"):
            c.iswap(int(a[0]), int(a[1]))
        elif g in ("rx", "ry", "rz"):
            getattr(c, g)(int(a[0]), theta=float(a[1]))
        elif g in ("fs", "fsim"):
            c.iswap(int(a[0]), int(a[1]), theta=-float(a[2]))
            c.cphase(int(a[0]), int(a[1]), theta=-float(a[3]))
        else:
            raise NotImplementedError(g)
    return c


def _register() -> None:  # abstractcircuit.py:242-373 _meta_apply
    def mk(g: str):
        def method(self: Circuit, *index: Any, **vars: Any) -> None:
            self._apply_named(g, *index, **vars)

        method.__name__ = g
        return method

    for g in Circuit.sgates + Circuit.vgates + ["diagonal"]:
        m = mk(g)
        setattr(Circuit, g, m)
        setattr(Circuit, g.upper(), m)
    for present, alias in Circuit.gate_aliases:
        setattr(Circuit, alias, getattr(Circuit, present))


_register()
