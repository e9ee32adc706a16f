#!/bin/bash
# GPU session 5 (round 2, 2 GPUs): fused / chunked final gather, 2-GPU bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/s5_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gather_check.py > $O/s5_gather.txt 2>&1
echo "rc=$?" >> $O/s5_gather.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/s5_bench2.txt 2> $O/s5_bench2_err.txt
echo done > $O/s5_done.txt
