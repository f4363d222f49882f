#!/bin/bash
# Round evidence on ONE B200 (run under gpurun): GPU tests, the bench line (+ per-stage roofline table, GEMM table, CPU
# baseline), the reference arm, the ncu launch list of the bench command and `--set full` captures of the top kernels.
# Usage: bash tools/collect_evidence.sh <tag>     (outputs under gpurun_out/)
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${tag}_stages.md CTCLIP_BENCH_GEMM_TABLE=gpurun_out/${tag}_gemm_table.txt \
  timeout 700 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> /dev/null; echo "ref rc=$?"
timeout 300 python bench.py --config zero_shot > gpurun_out/${tag}_bench_zero_shot.json 2> gpurun_out/${tag}_bench_zero_shot.err; echo "zero-shot rc=$?"
timeout 300 python tools/attn_tc_probe.py --reps 10 > gpurun_out/${tag}_attn_probe.txt 2>&1
timeout 300 python tools/peg_stream_probe.py > gpurun_out/${tag}_peg_probe.txt 2>&1
# launch list of the bench command (1 warm-up + 1 timed + 2 e2e steps; cold-cache serialised times: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stages > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
# --set full captures on a 1+1-layer step (same kernels and shapes, fewer launches)
CMD="python bench.py --steps 1 --warmup 1 --depth 1 --bert-layers 1 --no-cpu-baseline --no-stages"
for spec in "gemm_geglu:gemm_tc_kernel:8:1" "attn_tc_fwd:attn_tc_fwd_kernel:0:1" "attn_tc_bwd:attn_tc_bwd_kernel:0:1" "attn_dtab:attn_dtab_reduce:0:1" \
            "attn_short:attn_short:0:2" "peg_conv:peg_stream_conv_kernel:0:4" "peg_wgrad:peg_stream_wgrad_kernel:0:2" "ln_bwd:ln_bwd_kernel:2:1" \
            "geglu_bwd:geglu_bwd_kernel:0:1" "l2norm_bwd:l2norm_bwd_kernel:0:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c $cnt -f \
    -o gpurun_out/${tag}_prof_${name} $CMD > gpurun_out/${tag}_prof_${name}.log 2>&1
done
ls gpurun_out/${tag}_prof_*.ncu-rep
