#!/bin/bash
# GPU box: PEG parity tests, the stream-vs-general probe, one ncu capture of the plane-streaming kernels, a short bench
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -k peg -q -x 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_peg.txt
tail -8 gpurun_out/${TAG}_pytest_peg.txt
timeout 300 python tools/peg_stream_probe.py > gpurun_out/${TAG}_peg_probe.txt 2>&1
cat gpurun_out/${TAG}_peg_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:peg_stream -c 3 -o gpurun_out/${TAG}_ncu_peg -f \
    python tools/peg_stream_probe.py --quick > gpurun_out/${TAG}_ncu_peg.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_peg.log
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${TAG}_stages.md python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 500 gpurun_out/${TAG}_bench.json; echo; grep -i peg gpurun_out/${TAG}_stages.md
