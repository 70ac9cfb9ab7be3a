#!/bin/bash
# 2-GPU round: swap wire rate over NCCL send/recv vs peer-memory stores, multi-GPU parity tests, the N=2 bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
CHUNKS=26 TCB_SWAP_P2P=0 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n2_nccl.txt 2>&1
CHUNKS=26,28 TCB_SWAP_P2P=1 timeout 300 $TR tools/swap_bench.py > gpurun_out/swap_n2_p2p.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r2b.json 2> gpurun_out/bench_n2_r2b.err
grep -h "swap\|unavailable" gpurun_out/swap_n2_*.txt
tail -3 gpurun_out/pytest_n2.log
tail -c 1500 gpurun_out/bench_n2_r2b.json
