#!/bin/bash
# GPU box: the full -m gpu suite and a short bench with the stage table
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.txt
tail -6 gpurun_out/${TAG}_pytest_gpu.txt
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md CTCLIP_BENCH_GEMM_TABLE=gpurun_out/${TAG}_gemm_table.txt \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 400 gpurun_out/${TAG}_bench.json; echo; head -34 gpurun_out/${TAG}_stages.md
