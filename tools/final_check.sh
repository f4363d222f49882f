#!/bin/bash
# configs[4] geometry (dim 768, 24+24 layers, 512x512x320, patch (16,16,8)) at the batch size that fits 180 GB without recompute
mkdir -p gpurun_out
timeout 420 python bench.py --dim 768 --image 512 --frames 320 --depth 24 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r2v_bench_cfg4_b2.json 2> gpurun_out/r2v_bench_cfg4_b2.err
echo "rc=$?"; head -c 700 gpurun_out/r2v_bench_cfg4_b2.json; echo; tail -5 gpurun_out/r2v_bench_cfg4_b2.err | cut -c1-300
nvidia-smi --query-gpu=memory.used --format=csv
