#!/bin/bash
# 2-GPU box: NCCL data-parallel parity script (full per-tensor report + the peer-memory latent exchange stress) and bench.py at
# N = 1 and N = 2 on the same box
TAG=${1:-r2}
mkdir -p gpurun_out
DP_CHECK_REPORT=gpurun_out/${TAG}_dp_check_grads.txt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29612 tests/dp_check_multigpu.py > gpurun_out/${TAG}_dp_check.txt 2>&1
echo "dp_check rc=$?"; grep -E "dp_check|DP_CHECK|grad diff|rror|warn" gpurun_out/${TAG}_dp_check.txt | cut -c1-220 | head -30
timeout 600 python -m pytest tests/test_dp_multigpu.py -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench n1 rc=$?"; head -c 330 gpurun_out/${TAG}_bench_n1.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "bench n2 rc=$?"; head -c 330 gpurun_out/${TAG}_bench_n2.json; echo; grep -i -E "warn|error|exchange" gpurun_out/${TAG}_bench_n2.err | head -5
