"""Sliced tensor-network contraction behind DistributedContractor (configs[4])."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build, random_layers  # noqa: E402

pytestmark = pytest.mark.gpu


def _nodes_fn(params):
    import tensorcircuit_ng_b200 as tc

    c = tc.Circuit(4)
    c.rx(range(4), theta=params["x"])
    c.cnot([0, 1, 2], [1, 2, 3])
    c.ry(range(4), theta=params["y"])
    return c.expectation_before([tc.gates.z(), [-1]], reuse=False)


def test_value_and_grad_matches_expectation_ps(cuda):
    """The reference's own test, tests/test_miscs.py:275-304 (value == expectation_ps, atol 1e-6)."""
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    params = {"x": np.ones([4], dtype=np.float32), "y": 0.3 * np.ones([4], dtype=np.float32)}
    dc = DistributedContractor(_nodes_fn, params, {"slicing_reconf_opts": {"target_size": 2**2}, "max_repeats": 8})
    assert dc.nslices >= 2
    value, grad = dc.value_and_grad(params)
    assert grad["y"].shape == (4,)
    c = tc.Circuit(4)
    c.rx(range(4), theta=torch.ones(4))
    c.cnot([0, 1, 2], [1, 2, 3])
    c.ry(range(4), theta=0.3 * torch.ones(4))
    base = c.expectation_ps(z=[-1])
    np.testing.assert_allclose(float(value), float(base.real), atol=1e-6)
    v2 = dc.value(params)
    np.testing.assert_allclose(complex(v2), complex(base), atol=1e-6)
    # gradient against central differences of the statevector path
    eps = 1e-2
    for key in ("x", "y"):
        for i in (0, 3):
            vals = []
            for sgn in (+1, -1):
                p = {k: torch.tensor(v) for k, v in params.items()}
                p[key] = p[key].clone()
                p[key][i] += sgn * eps
                c = tc.Circuit(4)
                c.rx(range(4), theta=p["x"].cuda())
                c.cnot([0, 1, 2], [1, 2, 3])
                c.ry(range(4), theta=p["y"].cuda())
                vals.append(float(c.expectation_ps(z=[-1]).real))
            fd = (vals[0] - vals[1]) / (2 * eps)
            assert abs(float(grad[key][i]) - fd) <= 2e-3


@pytest.mark.parametrize("n,seed,target", [(8, 0, 2**4), (10, 1, 2**5), (12, 2, 2**30)])
def test_sliced_amplitudes_match_oracle(cuda, n, seed, target):
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tc_oracle

    ops = random_layers(n, 3, seed)
    bits = "".join(str((3 * i + seed) % 2) for i in range(n))

    def nodes_fn(_):
        return build(tc, n, ops).amplitude_before(bits)

    dc = DistributedContractor(nodes_fn, torch.zeros(1), {"slicing_reconf_opts": {"target_size": target}})
    if target < 2**20:
        assert dc.nslices >= 2
    amp = dc.value(torch.zeros(1))
    ref = build(tc_oracle, n, ops).amplitude(bits)
    assert abs(complex(amp) - complex(ref)) <= 1e-5
    # a plan without the diagonal->hyperedge rewrite (what a cotengra plan made elsewhere looks like)
    dc2 = DistributedContractor(nodes_fn, torch.zeros(1), {"slicing_reconf_opts": {"target_size": target}, "hyper_diagonal": False})
    assert abs(complex(dc2.value(torch.zeros(1))) - complex(ref)) <= 1e-5


def _nccl_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    torch.set_default_device(dev)
    params = {"x": np.ones([4], dtype=np.float32), "y": 0.3 * np.ones([4], dtype=np.float32)}
    dc = DistributedContractor(_nodes_fn, params, {"slicing_reconf_opts": {"target_size": 2**2}})
    v, g = dc.value_and_grad(params)
    torch.save({"v": v.cpu(), "gy": g["y"].cpu(), "mine": list(dc._my_slices()), "nslices": dc.nslices}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_distributed_contractor_two_gpus(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from tensorcircuit_ng_b200.experimental import DistributedContractor

    port = 29700 + (os.getpid() % 2000)
    out = str(tmp_path / "dc")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    parts = [torch.load(f"{out}.{r}", weights_only=False) for r in range(2)]
    assert sorted(parts[0]["mine"] + parts[1]["mine"]) == list(range(parts[0]["nslices"]))
    params = {"x": np.ones([4], dtype=np.float32), "y": 0.3 * np.ones([4], dtype=np.float32)}
    v1, g1 = DistributedContractor(_nodes_fn, params, {"slicing_reconf_opts": {"target_size": 2**2}}).value_and_grad(params)
    for p in parts:
        np.testing.assert_allclose(float(p["v"]), float(v1), atol=1e-6)
        np.testing.assert_allclose(p["gy"].numpy(), g1["y"].cpu().numpy(), atol=1e-5)


def test_rcs_amplitude_tn_vs_statevector(cuda):
    """configs[4] family at a size the statevector path can check: 5x5 depth-12 random circuit, amplitude of
    |0...0> by sliced tensor-network contraction (committed plan file) == wavefunction()[0]."""
    import pickle

    sys.path.insert(0, ROOT)
    import bench
    import tensorcircuit_ng_b200 as tc
    from tensorcircuit_ng_b200.experimental import DistributedContractor

    rows = cols = 5
    depth = 12
    c = bench.build_rcs(tc, rows, cols, depth)
    psi0 = complex(c.wavefunction()[0])
    nodes_fn = lambda _: bench.build_rcs(tc, rows, cols, depth).amplitude_before("0" * 25)  # noqa: E731
    td = pickle.load(open(bench.rcs_plan_path(rows, cols, depth, 30), "rb"))
    dc = DistributedContractor(nodes_fn, torch.zeros(1), tree_data=td)
    amp = complex(dc.value(torch.zeros(1)))
    assert abs(amp - psi0) <= 1e-6, (amp, psi0)
    # forced slicing of the same network
    dc2 = DistributedContractor(nodes_fn, torch.zeros(1), {"slicing_reconf_opts": {"target_size": 2**16}})
    assert dc2.nslices >= 2
    assert abs(complex(dc2.value(torch.zeros(1))) - psi0) <= 1e-6
