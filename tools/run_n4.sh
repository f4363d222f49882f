#!/bin/bash
# 4-GPU box: the data-parallel parity script at world 4 (peer-memory exchange with 4 ranks) and bench.py --gpus 4
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tests/dp_check_multigpu.py \
  > gpurun_out/${TAG}_dp_check_n4.txt 2>&1
echo "dp_check rc=$?"; grep -E "dp_check world|DP_CHECK|rror|warn" gpurun_out/${TAG}_dp_check_n4.txt | cut -c1-220 | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 4 --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err
echo "bench n4 rc=$?"; head -c 330 gpurun_out/${TAG}_bench_n4.json; echo; grep -i -E "warn|error|exchange" gpurun_out/${TAG}_bench_n4.err | head -5
