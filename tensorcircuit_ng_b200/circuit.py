"""
`Circuit` on the torch backend — host-side mirror of the reference's circuit front end for
the hot path (same method names, argument meaning and error behaviour):

  /root/reference/tensorcircuit/circuit.py:44-131 (init), :701-721 (wavefunction), :833-913
  /root/reference/tensorcircuit/basecircuit.py:52-66, :151-181, :183-371, :375-447, :562-640
  /root/reference/tensorcircuit/abstractcircuit.py:37-81, :114-373, :1523-1603

A gate is never "applied" here either: `apply_general_gate` appends a node and rewires the
front (basecircuit.py:277-293); numbers are produced when `wavefunction / expectation /
amplitude` hand the node list to the module-global `contractor` (cons.py) — which is the
drop-in boundary the B200 engine sits behind.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import cons, gates, tn
from .cons import contractor  # rebound by cons._set_global_contractor
from .gates import Gate

Tensor = Any


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


def _is_sequence(x: Any) -> bool:
    return isinstance(x, (list, tuple, range, np.ndarray))


# abstractcircuit.py:37-81
sgates = (
    ["i", "x", "y", "z", "h", "t", "s", "td", "sd", "wroot"]
    + ["cnot", "cz", "swap", "cy", "ox", "oy", "oz"]
    + ["toffoli", "fredkin"]
)
vgates = [
    "r", "cr", "u", "cu", "rx", "ry", "rz", "phase", "rxx", "ryy", "rzz", "cphase", "crx", "cry", "crz",
    "orx", "ory", "orz", "iswap", "any", "exp", "exp1",
]  # fmt: skip
diaggates = ["diagonal"]
gate_aliases = [["cnot", "cx"], ["fredkin", "cswap"], ["toffoli", "ccnot"], ["toffoli", "ccx"],
                ["any", "unitary"], ["sd", "sdg"], ["td", "tdg"]]  # fmt: skip


class Circuit:
    is_dm = False
    is_mps = False
    sgates, vgates, diaggates, gate_aliases = sgates, vgates, diaggates, gate_aliases

    def __init__(self, nqubits: int, inputs: Optional[Tensor] = None, **unsupported: Any) -> None:
        for k, v in unsupported.items():
            if v is not None:
                raise NotImplementedError(
                    f"Circuit({k}=...) is outside the B200 hot-path scope (SURVEY §2.1: MPS inputs, "
                    "SVD splitting and qudits are out of scope)"
                )
        self._d = 2
        self._nqubits = nqubits
        self.inputs = inputs
        self.circuit_param = {"nqubits": nqubits, "inputs": inputs}
        if inputs is None:
            nodes = self.all_zero_nodes(nqubits)
            self._front = [n.get_edge(0) for n in nodes]
        else:
            t = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(np.asarray(inputs), device=gates._device())
            t = t.to(torch.complex64).reshape(-1)
            N = t.shape[0]
            n = int(round(np.log2(N)))
            if (1 << n) != N or (n != nqubits and n != 2 * nqubits):
                raise ValueError(
                    f"inputs has {N} elements => {n} sites (dim=2), "
                    f"expected {nqubits} (state) or {2 * nqubits} (density matrix)"
                )
            node = Gate(t.reshape([2] * n))
            nodes = [node]
            self._front = [node.get_edge(i) for i in range(n)]
        self.coloring_nodes(nodes, flag="inputs")
        self._nodes: List[tn.Node] = nodes
        self._start_index = len(nodes)
        self._qir: List[Dict[str, Any]] = []
        self._extra_qir: List[Dict[str, Any]] = []
        self.state_tensor: Optional[tn.Node] = None

    # basecircuit.py:52-66 (the |0> leaves are host constants here; the engine never reads them)
    @staticmethod
    def all_zero_nodes(n: int, prefix: str = "qb-") -> List[tn.Node]:
        zero = gates._const(_ZERO)
        return [tn.Node(zero, name=prefix + str(x)) for x in range(n)]

    @staticmethod
    def coloring_nodes(nodes: Sequence[tn.Node], is_dagger: bool = False, flag: str = "inputs") -> None:
        for node in nodes:
            node.is_dagger = is_dagger  # type: ignore[attr-defined]
            node.flag = flag  # type: ignore[attr-defined]
            node.id = id(node)  # type: ignore[attr-defined]

    @staticmethod
    def copy_nodes(nodes: Sequence[tn.Node], dangling: Optional[Sequence[tn.Edge]] = None,
                   conj: Optional[bool] = False) -> Tuple[List[tn.Node], List[tn.Edge]]:  # fmt: skip
        ndict, edict = tn.copy(nodes, conjugate=bool(conj))
        newnodes = []
        for n in nodes:
            newn = ndict[n]
            newn.is_dagger = conj  # type: ignore[attr-defined]
            newn.flag = getattr(n, "flag", "") + "copy"  # type: ignore[attr-defined]
            newn.id = getattr(n, "id", id(n))  # type: ignore[attr-defined]
            newnodes.append(newn)
        newfront = []
        if not dangling:
            dangling = []
            for n in nodes:
                dangling.extend([e for e in n])  # type: ignore[union-attr]
        for e in dangling:
            newfront.append(edict[e])
        return newnodes, newfront

    def _copy(self, conj: Optional[bool] = False) -> Tuple[List[tn.Node], List[tn.Edge]]:
        return self.copy_nodes(self._nodes, self._front, conj)

    # basecircuit.py:183-371 ------------------------------------------------------------------
    def apply_general_gate(self, gate: Gate, *index: int, name: Optional[str] = None,
                           split: Optional[Dict[str, Any]] = None, mpo: bool = False,
                           diagonal: bool = False, ir_dict: Optional[Dict[str, Any]] = None) -> None:  # fmt: skip
        if name is None:
            name = ""
        if split is not None or mpo:
            raise NotImplementedError("split / mpo gates are outside the B200 hot-path scope (SURVEY §2.1)")
        gate_dict = {"gate": gate, "index": index, "name": name, "split": split, "mpo": mpo, "diagonal": diagonal}
        if ir_dict is not None:
            ir_dict.update(gate_dict)
        else:
            ir_dict = gate_dict
        self._qir.append(ir_dict)
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
            gate.name = name
            self._nodes.append(gate)
            for i, ind in enumerate(index):
                cn = tn.CopyNode(3, self._d, name=f"{name}_copy_{i}", device=gate.tensor.device)
                self.coloring_nodes([cn], flag="gate")
                self._nodes.append(cn)
                cn[0] ^ self._front[ind]
                cn[1] ^ gate[i]
                self._front[ind] = cn[2]
        self.state_tensor = None

    apply = apply_general_gate

    # abstractcircuit.py:114-240 ----------------------------------------------------------------
    def _apply_named(self, gname: str, *index: Any, **vars: Any) -> None:
        if isinstance(index[0], (int, np.integer)):
            self._apply_one(gname, *index, **vars)
        elif _is_sequence(index[0]):
            for i, ind in enumerate(zip(*index)):
                nvars = {}
                for k, v in vars.items():
                    try:
                        nvars[k] = v[i]
                    except Exception:  # pylint: disable=broad-except
                        nvars[k] = v
                self._apply_one(gname, *ind, **nvars)
        else:
            raise ValueError(
                f"Illegal index specification: each positional index must be an int "
                f"or a sequence/range of ints, got first element of type "
                f"{type(index[0]).__name__}: {index[0]!r}"
            )

    def _apply_one(self, gname: str, *index: int, **vars: Any) -> None:
        localname = vars.pop("name", gname)
        split = vars.pop("split", None)
        index = tuple(int(i) for i in index)
        if gname in sgates:
            gate = getattr(gates, gname)()
            self.apply_general_gate(gate, *index, name=localname, split=split, ir_dict={"gatef": getattr(gates, gname)})
            return
        if gname in diaggates:
            vars.setdefault("dim", self._d)
            gate = gates.diagonal_gate(**vars)
            self.apply_general_gate(gate, *index, name=localname, diagonal=True,
                                    ir_dict={"gatef": gates.diagonal_gate, "parameters": vars})  # fmt: skip
            return
        gatef = getattr(gates, gname + "_gate")
        gate = gates.memoised_gate(gatef, vars)
        self.apply_general_gate(gate, *index, name=localname, split=split,
                                ir_dict={"gatef": gatef, "parameters": vars})  # fmt: skip

    # circuit.py:701-721 -------------------------------------------------------------------------
    def wavefunction(self, form: str = "default") -> torch.Tensor:
        nodes, d_edges = self._copy()
        t = contractor(nodes, output_edge_order=d_edges)
        shape = {"default": [-1], "ket": [-1, 1], "bra": [1, -1]}[form]
        return t.tensor.reshape(shape)

    state = wavefunction

    def matrix(self) -> torch.Tensor:  # circuit.py:743-769
        n = self._nqubits
        eye = torch.eye(2**n, dtype=torch.complex64, device=gates._device())
        c = Circuit(n, inputs=eye)
        for d in self._qir:
            g = Gate(d["gate"].tensor)
            if hasattr(d["gate"], "_b200_kind"):
                g._b200_kind = d["gate"]._b200_kind  # type: ignore[attr-defined]
            c.apply_general_gate(g, *d["index"], name=d["name"], diagonal=d["diagonal"])
        return c.state().reshape(2**n, 2**n)

    # basecircuit.py:375-391 ---------------------------------------------------------------------
    def _copy_state_tensor(self, conj: bool = False, reuse: bool = True) -> Tuple[List[tn.Node], List[tn.Edge]]:
        if reuse:
            t = getattr(self, "state_tensor", None)
            if t is None:
                nodes, d_edges = self._copy()
                t = contractor(nodes, output_edge_order=d_edges)
                setattr(self, "state_tensor", t)
            ndict, edict = tn.copy([t], conjugate=conj)
            return [ndict[t]], [edict[e] for e in t.edges]
        return self._copy(conj)

    # basecircuit.py:393-447 ---------------------------------------------------------------------
    def expectation_before(self, *ops: Tuple[Any, Any], reuse: bool = True, **kws: Any) -> List[tn.Node]:
        nq = self._nqubits
        nodes1, edge1 = self._copy_state_tensor(reuse=reuse)
        nodes2, edge2 = self._copy_state_tensor(conj=True, reuse=reuse)
        nodes = nodes1 + nodes2
        newdang = edge1 + edge2
        occupied = set()
        for op, index in ops:
            if not isinstance(op, tn.Node):
                op = gates.num_to_tensor(op)
                op = Gate(gates._reshape2(op))
            else:
                op.tensor = op.tensor.to(torch.complex64)
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

    # circuit.py:833-913 -------------------------------------------------------------------------
    def expectation(self, *ops: Tuple[Any, Any], reuse: bool = True, enable_lightcone: bool = False,
                    noise_conf: Optional[Any] = None, **kws: Any) -> torch.Tensor:  # fmt: skip
        if noise_conf is not None:
            raise NotImplementedError("noisy expectation is outside the B200 hot-path scope (SURVEY §2.1)")
        if enable_lightcone:
            reuse = False
        nodes1 = self.expectation_before(*ops, reuse=reuse)
        if enable_lightcone:
            from .simplify import _full_light_cone_cancel

            nodes1 = _full_light_cone_cancel(nodes1)
        return contractor(nodes1).tensor

    # abstractcircuit.py:1523-1603 -----------------------------------------------------------------
    def expectation_ps(self, x: Optional[Sequence[int]] = None, y: Optional[Sequence[int]] = None,
                       z: Optional[Sequence[int]] = None, ps: Optional[Sequence[int]] = None,
                       reuse: bool = True, **kws: Any) -> torch.Tensor:  # fmt: skip
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

    # basecircuit.py:562-624 -----------------------------------------------------------------------
    def amplitude_before(self, l: Union[str, Tensor]) -> List[tn.Node]:
        no, d_edges = self._copy()
        if isinstance(l, str):
            l = [int(ch) for ch in l]
        lt = l if isinstance(l, torch.Tensor) else torch.as_tensor(np.asarray(l), device=gates._device())
        # quantum.py:166-183 onehot_d_tensor: the bitstring may be a runtime tensor
        endns = torch.nn.functional.one_hot(lt.to(torch.int64), 2).to(torch.complex64)
        ms = []
        for i in range(self._nqubits):
            n = tn.Node(endns[i])
            self.coloring_nodes([n], flag="measurement")
            ms.append(n)
            d_edges[i] ^ n.get_edge(0)
        no.extend(ms)
        return no

    def amplitude(self, l: Union[str, Tensor]) -> torch.Tensor:
        no = self.amplitude_before(l)
        return contractor(no).tensor

    def probability(self) -> torch.Tensor:  # basecircuit.py:626-640
        s = self.state()
        return (s.abs() ** 2).real

    def to_qir(self) -> List[Dict[str, Any]]:
        return self._qir


_ZERO = np.array([1.0, 0.0])


def _register() -> None:  # abstractcircuit.py:242-373 `_meta_apply`
    def mk(g: str):
        def method(self: Circuit, *index: Any, **vars: Any) -> None:
            self._apply_named(g, *index, **vars)

        method.__name__ = g
        method.__doc__ = f"Apply **{g.upper()}** gate on the circuit (tensorcircuit.gates.{g}_gate)."
        return method

    for g in sgates + vgates + diaggates:
        m = mk(g)
        setattr(Circuit, g, m)
        setattr(Circuit, g.upper(), m)
    for present, alias in gate_aliases:
        setattr(Circuit, alias, getattr(Circuit, present))


_register()


def expectation(*ops: Tuple[tn.Node, List[int]], ket: Tensor, bra: Optional[Tensor] = None,
                conj: bool = True, normalization: bool = False) -> torch.Tensor:  # fmt: skip
    """Module-level expectation <bra|ops|ket> for explicit state tensors (circuit.py:920-1065);
    only the bra == ket case is on the hot path."""
    if bra is not None:
        raise NotImplementedError("expectation(bra != ket) is outside the B200 hot-path scope")
    t = ket.tensor if isinstance(ket, tn.Node) else ket
    n = int(round(np.log2(t.numel())))
    c = Circuit(n, inputs=t)
    val = c.expectation(*ops)
    if normalization:
        from . import expect

        val = val / expect.operator_expectation(t.reshape(-1), n, [])
    return val
