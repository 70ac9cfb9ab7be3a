#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CHUNKS=26 TCB_SWAP_P2P=1 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n${N}_p2p_c.txt 2>&1
grep -h "swap\|unavailable" gpurun_out/swap_n${N}_p2p_c.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -2
