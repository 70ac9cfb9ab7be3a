"""Developer tool (torchrun, N ranks): wire rate of the global<->local qubit swap for chunk sizes / NCCL settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from tensorcircuit_ng_b200 import sharded
g = world.bit_length() - 1
n = 30 + g
for log_chunk in [int(x) for x in os.environ.get("CHUNKS", "26,27,28").split(",")]:
    sv = sharded.ShardedStatevector(n, sharded.TorchDistComm(), sharded.CudaExecutor(dev), chunk_elems=1 << log_chunk)
    pairs1 = [(n - 1, n - g - 1)]                       # one global qubit <-> the top local one
    pairsm = [(n - 1 - j, n - g - 1 - j) for j in range(g)]  # all global qubits at once
    for name, pairs in (("1 qubit", pairs1), (f"{g} qubits", pairsm)):
        if name != "1 qubit" and g == 1:
            continue
        sv.swap(pairs); torch.cuda.synchronize(); dist.barrier()
        s0 = sv.bytes_sent
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sv.swap(pairs)
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        ms = e0.elapsed_time(e1) / 3
        if rank == 0:
            print(f"chunk 2^{log_chunk} swap {name}: {ms:.2f} ms  {(sv.bytes_sent - s0) / 3 / ms / 1e6:.0f} GB/s per GPU per direction", flush=True)
    del sv; torch.cuda.empty_cache()
dist.destroy_process_group()
