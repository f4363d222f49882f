#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/peg_debug.py 2 26 24 24 15 > gpurun_out/peg_debug.txt 2>&1
timeout 400 python tools/peg_debug.py 8 24 24 24 10 >> gpurun_out/peg_debug.txt 2>&1
grep -E "runs with bad|bad;|rror|trap" gpurun_out/peg_debug.txt | head -40 | cut -c1-200
timeout 300 python tools/peg_stream_probe.py > gpurun_out/peg_probe_dbg.txt 2>&1
grep -E "stream|diff" gpurun_out/peg_probe_dbg.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -k peg -q 2>&1 | tail -3
