#!/bin/bash
# GPU session 7 (round 2, 8 GPUs): fused final gather and the bench line at N=8.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/gather_check.py > $O/s7_gather8.txt 2>&1
echo "rc=$?" >> $O/s7_gather8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 > $O/s7_bench8.txt 2> $O/s7_bench8_err.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/gather_check.py > $O/s7_gather4.txt 2>&1
echo done > $O/s7_done.txt
