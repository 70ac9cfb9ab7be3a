"""Sharded statevector on real GPUs (NCCL): needs >= 2 devices, otherwise only the single-GPU pieces run."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build, random_layers  # noqa: E402

pytestmark = pytest.mark.gpu


def test_pack_bits_kernel(cuda):
    from tensorcircuit_ng_b200 import _lib

    rng = np.random.default_rng(0)
    n = 14
    psi = (rng.normal(size=2**n) + 1j * rng.normal(size=2**n)).astype(np.complex64)
    for sel, pattern in [([0], 1), ([3], 0), ([2, 9], 2), ([1, 5, 13], 5), ([4, 6], 3)]:
        st = torch.from_numpy(psi).cuda()
        m = len(sel)
        block = 2 ** (n - m)
        idx = np.arange(block, dtype=np.int64)
        x = idx.copy()
        for p in sel:
            x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1))
        for k, p in enumerate(sel):
            x |= ((pattern >> k) & 1) << p
        for first, count in [(0, block), (block // 4, block // 2), (6, 10)]:
            buf = torch.zeros(count, dtype=torch.complex64, device="cuda")
            _lib.call("tcb_sv_pack_bits", st.data_ptr(), buf.data_ptr(), n, m, _lib.int_array(sel), pattern, first, count,
                      _lib.stream_ptr())  # fmt: skip
            assert np.array_equal(buf.cpu().numpy(), psi[x[first : first + count]])
            st2 = torch.zeros(2**n, dtype=torch.complex64, device="cuda")
            _lib.call("tcb_sv_unpack_bits", st2.data_ptr(), buf.data_ptr(), n, m, _lib.int_array(sel), pattern, first,
                      count, _lib.stream_ptr())  # fmt: skip
            want = np.zeros(2**n, dtype=np.complex64)
            want[x[first : first + count]] = psi[x[first : first + count]]
            assert np.array_equal(st2.cpu().numpy(), want)


def _nccl_worker(rank, world, port, n, seed, out, p2p):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["TCB_SWAP_P2P"] = str(p2p)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200 import sharded

    torch.set_default_device(dev)
    c = build(tc, n, random_layers(n, 3, seed))
    sv = sharded.evolve(c, chunk_elems=1 << 12)
    zz = sv.z_expectations([[0, n - 1], [2], [1, 3, 4]])
    amp = sv.amplitude([1, 0] * (n // 2))
    torch.save({"state": sv.state.cpu(), "pos_of": sv.pos_of, "zz": zz.cpu(), "amp": amp.cpu(), "swaps": sv.swaps_done,
                "norm": float(sv.norm2()[0]), "peer": bool(sv._peer)}, f"{out}.{rank}")  # fmt: skip
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nccl_matches_oracle(cuda, tmp_path, world, p2p):
    """p2p = 1: qubit swaps over peer memory (pack kernels store into the receiver's symmetric-memory staging
    buffer, several chunks per swap so both staging halves and the side-stream unpack are exercised);
    p2p = 0: the NCCL send/recv path."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from sharded_emu import gather_logical

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tc_oracle

    n, seed = 16, 11
    g = world.bit_length() - 1
    port = 29600 + (os.getpid() % 2000) + 7 * p2p
    out = str(tmp_path / "shard")
    mp.spawn(_nccl_worker, args=(world, port, n, seed, out, p2p), nprocs=world, join=True)
    parts = [torch.load(f"{out}.{r}", weights_only=False) for r in range(world)]
    assert parts[0]["swaps"] >= 1
    if not p2p:
        assert not any(p["peer"] for p in parts)
    ref = build(tc_oracle, n, random_layers(n, 3, seed)).wavefunction()
    full = gather_logical([p["state"].numpy() for p in parts], parts[0]["pos_of"], n, n - g)
    assert np.abs(full - ref).max() <= 1e-5
    p = np.abs(ref.astype(np.complex128)) ** 2
    idx = np.arange(2**n)

    def zexp(qs):
        s = np.ones(2**n)
        for q in qs:
            s *= 1 - 2 * ((idx >> (n - 1 - q)) & 1)
        return float(np.sum(p * s))

    want = [zexp([0, n - 1]), zexp([2]), zexp([1, 3, 4])]
    ai = int("10" * (n // 2), 2)
    for part in parts:
        assert np.allclose(part["zz"].numpy(), want, atol=1e-6)
        assert abs(complex(part["amp"][0]) - ref[ai]) <= 1e-5
        assert abs(part["norm"] - 1.0) <= 1e-5
