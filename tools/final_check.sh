#!/bin/bash
mkdir -p gpurun_out
timeout 420 python bench.py --dim 768 --image 512 --frames 320 --depth 24 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r2x_bench_cfg4_b2.json 2> gpurun_out/r2x_bench_cfg4_b2.err
echo "rc=$?"; head -c 300 gpurun_out/r2x_bench_cfg4_b2.json; echo
