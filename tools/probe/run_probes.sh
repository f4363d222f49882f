#!/bin/bash
# GPU box: bring-up probes for the tcgen05 attention kernels; every probe in its own process under a timeout.
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
OUT=../../gpurun_out/probes.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT 2>&1
for args in "ss64 64" "ss64 96" "ss64 192" "diag 0" "diag 1" "ts 96 0 48 96" "ts 64 0 32 64" "ts 128 0 64 128" "mna" "ns16"; do
  echo "== umma_probe $args" >> $OUT
  timeout 60 ./umma_probe $args >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
cat $OUT

