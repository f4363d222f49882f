#!/bin/bash
# GPU box: attention probe + ncu of the tcgen05 backward + the full -m gpu suite + a short bench with the stage table
TAG=${1:-r2d}
mkdir -p gpurun_out
bash tools/probe/run_probes.sh > /dev/null 2>&1; cp gpurun_out/probes.txt gpurun_out/${TAG}_probes.txt; grep -E "probe|rc=" gpurun_out/${TAG}_probes.txt
python tools/attn_tc_probe.py --reps 10 > gpurun_out/${TAG}_attn_probe.txt 2>&1
cat gpurun_out/${TAG}_attn_probe.txt
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 1 -c 1 -o gpurun_out/${TAG}_ncu_attn_bwd -f \
    python tools/attn_tc_probe.py --reps 1 --only tc --bwd-warps 108 > gpurun_out/${TAG}_ncu_bwd.log 2>&1
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.txt
tail -15 gpurun_out/${TAG}_pytest_gpu.txt
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md CTCLIP_BENCH_GEMM_TABLE=gpurun_out/${TAG}_gemm_table.txt \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json; echo; head -24 gpurun_out/${TAG}_stages.md
