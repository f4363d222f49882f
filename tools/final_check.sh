#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; head -c 330 gpurun_out/r2s_bench.json; echo
