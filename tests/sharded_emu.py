"""TEST INFRASTRUCTURE: a CPU executor for sharded.py built on the kernel-logic emulator (tests/emu) and
numpy, plus an in-process thread communicator.  Never imported by the package."""
import ctypes
import os
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emu():
    return ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libpass_emu.so"))


def _deposit(i, sel):
    x = i.copy()
    for p in sel:  # ascending
        x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1))
    return x


class EmuExecutor:
    def __init__(self, tile_bits=10, low_bits=3):
        self.opts = dict(tile_bits=tile_bits, low_bits=low_bits)
        self.emu = _emu()

    def zeros(self, n):
        return torch.zeros(n, dtype=torch.complex64)

    empty = zeros

    def set_one(self, state):
        state[0] = 1.0

    def init_product(self, state, nl, vecs, nq, index_base):
        v = vecs.numpy()
        x = np.arange(1 << nl, dtype=np.int64) | index_base
        out = np.ones(1 << nl, dtype=np.complex64)
        for p in range(nq):
            out *= v[p][(x >> p) & 1]
        state.numpy()[:] = out

    def run_gates(self, state, ops, gatebuf, nq, nl, pos_of, index_base, cache, key):
        from tensorcircuit_ng_b200 import passplan

        plan = cache.get(key)
        if plan is None:
            plan = passplan.compile_plan(list(ops), nq, nbits_local=nl, pos_of=pos_of, **self.opts)
            cache[key] = plan
        st = state.numpy()
        buf = gatebuf.numpy() if isinstance(gatebuf, torch.Tensor) else gatebuf
        for step in plan.steps:
            if isinstance(step, passplan.PassStep):
                prog = np.ascontiguousarray(step.program)
                rc = self.emu.emu_run_pass(
                    st.ctypes.data_as(ctypes.c_void_p), nl, ctypes.c_longlong(1),
                    prog.ctypes.data_as(ctypes.c_void_p), len(prog), step.tile_bits, step.low_bits,
                    buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(0), ctypes.c_ulonglong(index_base),
                )  # fmt: skip
                assert rc == 0, rc
            else:
                g, bp = step.gate, step.bitpos
                k = g.k
                idx = np.arange(1 << nl, dtype=np.int64) | index_base
                if g.is_diag:
                    d = buf[g.mat_off : g.mat_off + (2**k if g.kind[0] == "diagvec" else 4**k)]
                    if g.kind[0] != "diagvec":
                        d = np.diag(d.reshape(2**k, 2**k))
                    sel = np.zeros(1 << nl, dtype=np.int64)
                    for i, p in enumerate(bp):
                        sel |= ((idx >> p) & 1) << (k - 1 - i)
                    st *= d[sel]
                else:
                    assert all(p < nl for p in bp)
                    m = buf[g.mat_off : g.mat_off + 4**k].reshape([2] * (2 * k))
                    axes = [nl - 1 - p for p in bp]
                    psi = np.tensordot(m, st.reshape([2] * nl), axes=[list(range(k, 2 * k)), axes])
                    st[:] = np.ascontiguousarray(np.moveaxis(psi, list(range(k)), axes)).reshape(-1)

    def pack(self, state, buf, nl, sel, pattern, first, count, unpack):
        i = np.arange(first, first + count, dtype=np.int64)
        x = _deposit(i, sel)
        for k, p in enumerate(sel):
            x |= ((pattern >> k) & 1) << p
        if unpack:
            state.numpy()[x] = buf.numpy()[:count]
        else:
            buf.numpy()[:count] = state.numpy()[x]

    def expect_z(self, state, nl, masks, index_base):
        p = np.abs(state.numpy().astype(np.complex128)) ** 2
        idx = np.arange(1 << nl, dtype=np.int64) | index_base
        out = []
        for m in masks:
            par = np.zeros(1 << nl, dtype=np.int64)
            v = idx & m
            while np.any(v):
                par ^= v & 1
                v >>= 1
            out.append(float(np.sum(p * (1 - 2 * par))))
        return torch.tensor(out, dtype=torch.float64)

    def read(self, state, idx):
        return state[idx : idx + 1].clone()


class ThreadWorld:
    """world_size ranks as threads of one process."""

    def __init__(self, world):
        self.world = world
        self.cv = threading.Condition()
        self.mail = {}
        self.red = {}
        self.gen = 0

    def comm(self, rank):
        return _ThreadComm(self, rank)


class _ThreadComm:
    def __init__(self, w, rank):
        self.w, self.rank, self.world = w, rank, w.world
        self.seq = 0
        self.rseq = 0

    def exchange(self, sends, recvs):
        w = self.w
        with w.cv:
            for peer, t in sends:
                w.mail[(self.rank, peer, self.seq)] = t.clone()
            w.cv.notify_all()
            for peer, t in recvs:
                key = (peer, self.rank, self.seq)
                while key not in w.mail:
                    w.cv.wait(timeout=60)
                t.copy_(w.mail.pop(key))
        self.seq += 1

    def all_reduce_sum(self, t):
        w = self.w
        with w.cv:
            key = self.rseq
            acc = w.red.setdefault(key, [0, None, 0])
            acc[1] = t.clone() if acc[1] is None else acc[1] + t
            acc[0] += 1
            w.cv.notify_all()
            while w.red[key][0] < self.world:
                w.cv.wait(timeout=60)
            t.copy_(w.red[key][1])
            acc[2] += 1
            if acc[2] == self.world:
                del w.red[key]
        self.rseq += 1
        return t


def gather_logical(states, pos_of, n, nl):
    """Full state in canonical order (qubit 0 = MSB) from per-rank shards under layout pos_of."""
    full = np.concatenate([np.asarray(s) for s in states])  # physical index = rank << nl | local
    x = np.arange(1 << n, dtype=np.int64)  # logical index
    phys = np.zeros(1 << n, dtype=np.int64)
    for q in range(n):
        phys |= ((x >> (n - 1 - q)) & 1) << pos_of[q]
    return full[phys]
