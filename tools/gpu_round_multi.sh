#!/bin/bash
# N-GPU round (N = $1): swap wire rate over peer memory vs NCCL, multi-GPU parity tests, the driver's bench line
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CHUNKS=26 TCB_SWAP_P2P=1 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n${N}_p2p.txt 2>&1
CHUNKS=26 TCB_SWAP_P2P=0 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n${N}_nccl.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/pytest_n${N}.log 2>&1
timeout 1200 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_r2c.json 2> gpurun_out/bench_n${N}_r2c.err
grep -h "swap\|unavailable" gpurun_out/swap_n${N}_*.txt
tail -3 gpurun_out/pytest_n${N}.log
python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_r2c.json"):
    if l.startswith("{"):
        d = json.loads(l)
        c = d["config"]
        print("qaoa", d["value"], d["ms_per_step"], "swaps", c.get("swaps"), "passes", c.get("hbm_passes"), "swap_ms", c.get("swap_ms_per_step"), "nvlink", c.get("nvlink_gbs_per_gpu"), "local", c.get("local_ms_per_step"), "e2e", d["e2e"]["ms_per_step"], "parity", (d.get("parity") or {}).get("ok"))
        for k, v in d.get("sub_records", {}).items():
            print(k, {x: v.get(x) for x in ("value", "ms_per_step", "skipped", "error")}, (v.get("config") or {}).get("nvlink_gbs_per_gpu"), (v.get("config") or {}).get("swap_ms_per_step"), (v.get("config") or {}).get("local_ms_per_step"))
PY
tail -5 gpurun_out/bench_n${N}_r2c.err
