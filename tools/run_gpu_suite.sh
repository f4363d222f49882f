#!/bin/bash
# GPU box: full -m gpu test suite, then a short bench with the per-stage table. Outputs under gpurun_out/<tag>_*.
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/${TAG}_pytest_gpu.txt
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.txt
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md CTCLIP_BENCH_GEMM_TABLE=gpurun_out/${TAG}_gemm_table.txt \
  python bench.py --steps 5 --warmup 3 ${BENCH_FLAGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_pytest_gpu.txt; head -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
