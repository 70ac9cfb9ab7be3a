#!/bin/bash
# N-GPU check of the swap path only (N = $1): wire rate over peer memory, then the QAOA bench line without sub-records
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CHUNKS=26 TCB_SWAP_P2P=1 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n${N}_p2p_b.txt 2>&1
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-sub-records > gpurun_out/bench_n${N}_r2d.json 2> gpurun_out/bench_n${N}_r2d.err
grep -h "swap\|unavailable" gpurun_out/swap_n${N}_p2p_b.txt
python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_r2d.json"):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]
        print("qaoa", d["value"], d["ms_per_step"], "swaps", c.get("swaps"), "passes", c.get("hbm_passes"), "swap_ms", c.get("swap_ms_per_step"), "nvlink", c.get("nvlink_gbs_per_gpu"), "local", c.get("local_ms_per_step"), "e2e", d["e2e"]["ms_per_step"], "parity", (d.get("parity") or {}).get("ok"))
PY
tail -3 gpurun_out/bench_n${N}_r2d.err
