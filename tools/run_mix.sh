#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_ctclip_gpu.py tests/test_bert_gpu.py -q -s 2>&1 | grep -E "un-forced|passed|failed|Error|error" | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:peg_stream -c 3 -o gpurun_out/${TAG}_ncu_peg -f \
    python tools/peg_stream_probe.py --quick > gpurun_out/${TAG}_ncu_peg.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_peg.log
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/${TAG}_bench.json; echo; grep -E "geglu_bwd|l2norm_bwd|peg_|gelu_bwd" gpurun_out/${TAG}_stages.md
