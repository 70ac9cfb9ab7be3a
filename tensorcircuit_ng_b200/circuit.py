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

import math

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

    def __del__(self) -> None:
        # Node <-> Edge reference cycles: without this every dropped circuit (630 nodes, ~3000 edges for the QAOA
        # workload) waits for the cyclic collector, whose full collections cost ~200 ms in a process with torch
        # loaded (measured: +26 ms per step on average).  Cutting the node -> edge references lets reference
        # counting free the network at once.  (`_copy()` hands out copies, so nothing outside depends on these.)
        try:
            for nd in self._nodes:
                nd.edges = []
            self._front = []
        except Exception:  # pylint: disable=broad-except  (interpreter shutdown)
            pass

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
                try:  # which qubit pairs the circuit couples: a hint for the Z-moment table (cons._z_moment)
                    t.tensor._b200_pair_hint = {  # type: ignore[attr-defined]
                        (min(d["index"]), max(d["index"])) for d in self._qir if len(d["index"]) == 2
                    }
                except Exception:  # pylint: disable=broad-except  (wrapper tensors)
                    pass
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
        """abstractcircuit.py:375-414: a shallow copy of the instruction list."""
        return list(self._qir)

    # ---- QIR / JSON circuit input and output (abstractcircuit.py:417-496, 1249-1268, 1354-1390) ----------
    @classmethod
    def from_qir(cls, qir: List[Dict[str, Any]], circuit_params: Optional[Dict[str, Any]] = None,
                 allow_channel: bool = False) -> "Circuit":  # fmt: skip
        """Rebuild a circuit from its instruction list (`abstractcircuit.py:417-496`): entries with `parameters`
        re-evaluate `gatef(**parameters)`, the others call the fixed-gate factory."""
        circuit_params = dict(circuit_params or {})
        if "nqubits" not in circuit_params:
            circuit_params["nqubits"] = 1 + max((max(d["index"]) for d in qir), default=0)
        c = cls(**circuit_params)
        for d in qir:
            if d.get("is_channel", False):
                if allow_channel:
                    raise NotImplementedError("noise channels are outside the B200 hot-path scope (SURVEY §2.1)")
                continue
            if d.get("mpo", False) or d.get("split"):
                raise NotImplementedError("split / mpo gates are outside the B200 hot-path scope (SURVEY §2.1)")
            gatef = d["gatef"]
            if "parameters" not in d:
                c.apply_general_gate(gatef(), *d["index"], name=d["name"], ir_dict={"gatef": gatef})
            else:
                params = dict(d["parameters"])
                c.apply_general_gate(gatef(**params), *d["index"], name=d["name"], diagonal=d.get("diagonal", False),
                                     ir_dict={"gatef": gatef, "parameters": params})  # fmt: skip
        return c

    def to_json(self, file: Optional[str] = None, simplified: bool = False) -> Any:
        """`abstractcircuit.py:1249-1268`: the circuit as the reference's JSON list (`translation.qir2json`)."""
        import json

        tcqasm = qir2json(self.to_qir(), simplified=simplified)
        if file is not None:
            with open(file, "w") as f:
                json.dump(tcqasm, f)
        return json.dumps(tcqasm)

    @classmethod
    def from_json(cls, jsonstr: Any, circuit_params: Optional[Dict[str, Any]] = None) -> "Circuit":
        """`abstractcircuit.py:1354-1374`."""
        import json

        if isinstance(jsonstr, str):
            jsonstr = json.loads(jsonstr)
        return cls.from_qir(json2qir(jsonstr), circuit_params)

    @classmethod
    def from_json_file(cls, file: str, circuit_params: Optional[Dict[str, Any]] = None) -> "Circuit":
        """`abstractcircuit.py:1376-1390`."""
        import json

        with open(file, "r") as f:
            return cls.from_json(json.load(f), circuit_params)

    # basecircuit.py:626-640 ---------------------------------------------------------------------
    def probability(self) -> torch.Tensor:
        s = self.wavefunction().reshape(-1)
        return s.real**2 + s.imag**2

    def _sampler(self) -> Any:
        from . import sampling

        nodes, _ = self._copy_state_tensor()  # the cached state (contracted once per circuit)
        return sampling.StateSampler(nodes[0].tensor.reshape(-1), self._nqubits)

    # basecircuit.py:449-558 ---------------------------------------------------------------------
    def perfect_sampling(self, status: Optional[Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One bitstring ([n] float32 of 0/1) and its probability; `status`: n uniforms (qubit 0 first)."""
        return self.measure(*range(self._nqubits), with_prob=True, status=status)

    def measure(self, *index: int, with_prob: bool = False, status: Optional[Tensor] = None) -> Tuple[torch.Tensor, Any]:
        """`measure_jit` (`:461-558`): qubits are read in the order given, qubit `index[k]` gives 1 iff
        status[k] - P(0 | earlier outcomes) + eps > 0.  All qubits in natural order: one conditional walk
        over the resident state (`tcb_sv_sample` mode 1); a subset: the same rule on its marginal."""
        from . import sampling

        n = self._nqubits
        idx = [i if i >= 0 else n + i for i in index]
        sampler = self._sampler()
        dev = sampler.psi.device
        st = sampling.uniforms([len(idx)], None, dev) if status is None else torch.as_tensor(status).reshape(-1)
        if len(st) != len(idx):
            raise ValueError(f"status must have one entry per measured qubit ({len(idx)})")
        if idx == list(range(n)):
            flat, prob = sampler.draw(st.reshape(1, n), mode=1)
            bits = ((flat.reshape(1, 1) >> torch.arange(n - 1, -1, -1, device=dev)) & 1).reshape(n).to(torch.float32)
            return bits, (prob[0].to(torch.float32) if with_prob else -1.0)
        if len(set(idx)) != len(idx) or len(idx) > 24:
            raise ValueError("measure: qubits must be distinct (and at most 24 for a subset)")
        p = (sampler.psi.real.double() ** 2 + sampler.psi.imag.double() ** 2).reshape([2] * n)
        rest = [q for q in range(n) if q not in idx]
        marg = p.sum(dim=rest) if rest else p  # axes = the measured qubits in ascending order
        if len(idx) > 1:
            marg = marg.permute(*np.argsort(np.argsort(idx)).tolist())  # ... now in the order they are read
        marg = (marg / marg.sum()).cpu().numpy().reshape(-1)
        u = st.detach().cpu().numpy().astype(np.float64)
        lo, hi, out, pr = 0, len(marg), [], 1.0
        cum = np.concatenate([[0.0], np.cumsum(marg)])
        for k in range(len(idx)):
            mid = (lo + hi) // 2
            m0, m1 = cum[mid] - cum[lo], cum[hi] - cum[mid]
            p0 = m0 / (m0 + m1) if m0 + m1 > 0 else 1.0
            one = u[k] - p0 + 0.31415926e-12 > 0
            out.append(1.0 if one else 0.0)
            pr *= (1.0 - p0) if one else p0
            lo, hi = (mid, hi) if one else (lo, mid)
        bits = torch.tensor(out, dtype=torch.float32, device=dev)
        return bits, (torch.tensor(pr, dtype=torch.float32, device=dev) if with_prob else -1.0)

    measure_jit = measure

    # basecircuit.py:1401-1512 -------------------------------------------------------------------
    def sample(self, batch: Optional[int] = None, allow_state: bool = False, readout_error: Optional[Any] = None,
               format: Optional[str] = None, random_generator: Optional[Any] = None, status: Optional[Tensor] = None,
               jittable: bool = True, format_: Optional[str] = None) -> Any:  # fmt: skip
        """Batched sampling.  `allow_state=True`: CDF inversion with one uniform per shot
        (`probability_sample`); `allow_state=False`: the perfect-sampling rule with n uniforms per shot.
        Both read the resident state (one segment-mass pass + one segment per shot); formats as in
        `quantum.sample2all`."""
        from . import quantum, sampling

        if readout_error is not None:
            raise NotImplementedError("readout_error is outside the B200 hot-path scope (SURVEY §2.1)")
        format = format if format is not None else format_
        n = self._nqubits
        nbatch = 1 if batch is None else int(batch)
        sampler = self._sampler()
        dev = sampler.psi.device
        if allow_state:
            st = sampling.uniforms([nbatch], random_generator, dev) if status is None else torch.as_tensor(status)
            if st.reshape(-1).shape[0] != nbatch:
                raise ValueError(f"status must have shape [{nbatch}]")
            ch, prob = sampler.draw(st.reshape(nbatch), mode=0)
        else:
            st = sampling.uniforms([nbatch, n], random_generator, dev) if status is None else torch.as_tensor(status)
            st = st.reshape(1, n) if st.dim() == 1 else st
            if tuple(st.shape) != (nbatch, n):
                raise ValueError(f"status must have shape [{nbatch}, {n}]")
            ch, prob = sampler.draw(st, mode=1)
        if format is None:  # backward-compatible form: (configuration, probability) per shot
            confg = quantum.sample_int2bin(ch, n)
            if allow_state:
                r = list(zip(confg, prob.to(torch.float32)))
            else:
                r = [(c.to(torch.float32), p) for c, p in zip(confg, prob.to(torch.float32))]
            return r[0] if batch is None else r
        if n > 32:
            if format == "sample_bin":
                return quantum.sample_int2bin(ch, n)
            if format == "count_dict_bin":
                from collections import Counter

                return dict(Counter("".join(str(int(b)) for b in row) for row in quantum.sample_int2bin(ch, n).cpu().tolist()))
            raise ValueError(f"n={n} is too large for measurement representaion: {format}")
        return quantum.sample2all(ch, n, format=format, jittable=jittable)

    # abstractcircuit.py:1269-1351 ----------------------------------------------------------------
    # Google qsim text format (the public RCS circuit files that feed config 5): line 1 = qubit count,
    # then "<moment> <gate> <qubits...> [<params...>]".  gate -> (method, number of qubits, fixed
    # parameters, names of the trailing numeric parameters); `fs` / `fsim` expands to iswap(-theta) +
    # cphase(-phi) as in the reference (`:1340-1343`).
    _QSIM: Dict[str, Tuple[str, int, Dict[str, float], Tuple[str, ...]]] = {
        "h": ("h", 1, {}, ()), "x": ("x", 1, {}, ()), "y": ("y", 1, {}, ()), "z": ("z", 1, {}, ()),
        "s": ("phase", 1, {"theta": np.pi / 2}, ()), "t": ("phase", 1, {"theta": np.pi / 4}, ()),
        "x_1_2": ("rx", 1, {"theta": np.pi / 2}, ()), "y_1_2": ("ry", 1, {"theta": np.pi / 2}, ()),
        "z_1_2": ("rz", 1, {"theta": np.pi / 2}, ()),
        "w_1_2": ("u", 1, {"theta": np.pi / 2, "phi": -np.pi / 4, "lbd": np.pi / 4}, ()),
        "hz_1_2": ("wroot", 1, {}, ()),
        "cnot": ("cnot", 2, {}, ()), "cx": ("cx", 2, {}, ()), "cy": ("cy", 2, {}, ()), "cz": ("cz", 2, {}, ()),
        "is": ("iswap", 2, {}, ()), "iswap": ("iswap", 2, {}, ()),
        "rx": ("rx", 1, {}, ("theta",)), "ry": ("ry", 1, {}, ("theta",)), "rz": ("rz", 1, {}, ("theta",)),
    }  # fmt: skip

    @classmethod
    def from_qsim_file(cls, file: str, circuit_params: Optional[Dict[str, Any]] = None) -> "Circuit":
        with open(file, "r") as f:
            lines = f.readlines()
        params = dict(circuit_params or {})
        params.setdefault("nqubits", int(lines[0]))
        return cls._apply_qsim(cls(**params), lines)

    @staticmethod
    def _apply_qsim(c: "Circuit", qsim_str: Sequence[str]) -> "Circuit":
        for line in qsim_str[1:]:
            tok = line.strip().split()
            if not tok:
                continue
            name = tok[1].lower()
            if name in ("fs", "fsim"):
                i, j, theta, phi = int(tok[2]), int(tok[3]), float(tok[4]), float(tok[5])
                c.iswap(i, j, theta=-theta)
                c.cphase(i, j, theta=-phi)
                continue
            spec = Circuit._QSIM.get(name)
            if spec is None:
                raise NotImplementedError(f"qsim gate `{tok[1]}` is not supported")
            method, nq, fixed, pnames = spec
            kws: Dict[str, Any] = dict(fixed)
            for pn, v in zip(pnames, tok[2 + nq :]):
                kws[pn] = float(v)
            getattr(c, method)(*[int(t) for t in tok[2 : 2 + nq]], **kws)
        return c


_ZERO = np.array([1.0, 0.0])


# ---- translation.py:602-719 (tensor <-> JSON lists, qir <-> JSON dicts) ------------------------------
def tensor_to_json(a: Any) -> Any:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return [np.real(a).tolist(), np.imag(a).tolist()]
    return [np.real(a).tolist()]


def json_to_tensor(a: Any) -> Any:
    ar = np.array(a[0])
    if len(a) == 1:
        return ar
    assert len(a) == 2
    return ar + 1.0j * np.array(a[1])


def gate_json_name(gatef: Any) -> str:
    """The name the reference stores in JSON (`gatef.n`, translation.py:667): `rx`, `cnot`, `exp1`, `any` ..."""
    n = getattr(gatef, "n", None)
    if n is not None:
        return str(n)
    fn = getattr(gatef, "__name__", str(gatef))
    special = {"exponential_gate_unity": "exp1", "exponential_gate": "exp", "g": None}
    if fn in special and special[fn] is not None:
        return special[fn]  # type: ignore[return-value]
    for name in vgates + sgates + diaggates:
        if getattr(gates, name + "_gate", None) is gatef or getattr(gates, name, None) is gatef:
            return name
    return fn[:-5] if fn.endswith("_gate") else fn


def get_u_parameter(m: np.ndarray) -> Tuple[float, float, float]:
    """gates.py:606-627: (theta, phi, lambda) of the U gate equal to a 2x2 unitary up to a global phase."""
    u = np.linalg.det(m) ** (-1 / 2) * m
    theta = 2 * np.arccos(min(1.0, float(np.abs(u[1, 1]))))
    plus, minus = 2 * np.angle(u[1, 1]), -2 * np.angle(u[1, 0])
    return float(theta), float((plus - minus) / 2), float((plus + minus) / 2)


def qir2json(qir: List[Dict[str, Any]], simplified: bool = False) -> List[Dict[str, Any]]:
    """translation.py:631-690."""
    out = []
    for r in qir:
        if r.get("is_channel", False):
            continue
        t = r["gate"].tensor
        d = int(round(math.sqrt(t.numel())))
        nm = t.detach().cpu().numpy().reshape(d, d) if not r.get("diagonal", False) else np.diag(t.detach().cpu().numpy().reshape(-1))
        nmr, nmi = tensor_to_json(nm.astype(np.complex64))
        uparams = [float(p) for p in get_u_parameter(nm.astype(np.complex128))] if nm.shape == (2, 2) else []
        params = {k: tensor_to_json(v) for k, v in r.get("parameters", {}).items()}
        item: Dict[str, Any] = {"name": gate_json_name(r["gatef"]), "qubits": list(r["index"])}
        unsupported = ["any", "unitary", "mpo", "exp", "exp1", "r", "cr"]
        if not simplified:
            item.update({"matrix": [nmr, nmi], "uparams": uparams, "parameters": params, "mpo": bool(r.get("mpo", False))})
        else:
            if item["name"] in unsupported and uparams:
                item.update({"uparams": uparams})
            elif item["name"] in unsupported:
                item.update({"matrix": [nmr, nmi]})
            if params:
                item.update({"parameters": params})
        out.append(item)
    return out


def json2qir(tcqasm: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """translation.py:693-719."""
    qir = []
    for d in tcqasm:
        param = {k: json_to_tensor(v) for k, v in d.get("parameters", {}).items()}
        if param.get("dim") is not None:
            param["dim"] = int(np.asarray(param["dim"]))
        name = d["name"]
        gatef = getattr(gates, name + "_gate", None) or getattr(gates, name)
        if name == "diagonal":
            gatef = gates.diagonal_gate
        entry = {"index": tuple(d["qubits"]), "mpo": d.get("mpo", False), "split": None, "parameters": param,
                 "gatef": gatef, "name": name, "diagonal": name in diaggates}  # fmt: skip
        if not param and name in sgates:
            entry.pop("parameters")
        qir.append(entry)
    return qir


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
