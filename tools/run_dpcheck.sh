#!/bin/bash
# 2-GPU box: the NCCL data-parallel parity script with its full per-tensor report
TAG=${1:-r2k}
mkdir -p gpurun_out
DP_CHECK_REPORT=gpurun_out/${TAG}_dp_check_grads.txt python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29612 tests/dp_check_multigpu.py > gpurun_out/${TAG}_dp_check.txt 2>&1
grep -E "dp_check|DP_CHECK|grad diff" gpurun_out/${TAG}_dp_check.txt | cut -c1-220
head -40 gpurun_out/${TAG}_dp_check_grads.txt | cut -c1-200
