#!/bin/bash
# 2-GPU box: NCCL data-parallel parity test, the attention probe, and bench.py at N = 1 and N = 2
TAG=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests/test_dp_multigpu.py -q -s 2>&1 | grep -v '^E  ' | tail -60 > gpurun_out/${TAG}_dp_multigpu.txt
grep -E 'dp_check|grad diff|DP_CHECK|passed|failed' gpurun_out/${TAG}_dp_multigpu.txt | head -30
python tools/attn_tc_probe.py --reps 10 > gpurun_out/${TAG}_attn_probe.txt 2>&1
cat gpurun_out/${TAG}_attn_probe.txt
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench n1 rc=$?"; head -c 400 gpurun_out/${TAG}_bench_n1.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "bench n2 rc=$?"; head -c 400 gpurun_out/${TAG}_bench_n2.json; echo; tail -5 gpurun_out/${TAG}_bench_n2.err
