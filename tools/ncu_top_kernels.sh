#!/bin/bash
# One `ncu --set full` capture per hot kernel on a 1+1-layer, bs8 full-resolution step (same kernels and shapes as the
# 12+12 bench; fewer launches). Usage (on the GPU box): bash tools/ncu_top_kernels.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --depth 1 --bert-layers 1 --no-cpu-baseline"
# name:kernel-base-name:skip:count. Launch order inside a step: BERT (1 layer) first, then the spatial layer, then the
# temporal layer; so launch #1 of the attention kernels is the spatial one, and GEMM launches 8/9 are FF1 (GEGLU) / FF2.
for spec in "attn_fwd:attn_fwd_kernel:1:1" "attn_dkv:attn_bwd_dkv_kernel:1:1" "attn_dq:attn_bwd_dq_kernel:1:1" "peg:peg_conv2_kernel:0:2" "gemm:gemm_tc_kernel:5:5"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c $cnt -f -o gpurun_out/prof_${tag}_${name} $CMD > gpurun_out/prof_${tag}_${name}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
