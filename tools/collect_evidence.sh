#!/bin/bash
# Round evidence on ONE B200 (run under gpurun): GPU tests, smoke, the bench line (+ per-stage roofline table, GEMM table, CPU
# baseline), the reference arm, the zero-shot config, the isolated probes, the ncu launch list of the bench command and
# `--set full` captures of the top kernels. gpurun copies back at most 64 MiB of gpurun_out/: the .ncu-rep files stay in /tmp on
# the box, only their text summaries (tools/ncu_summary.py) are kept.
# Usage: bash tools/collect_evidence.sh <tag> [quick]
tag=${1:-r2}
mkdir -p gpurun_out /tmp/ncu_$tag
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${tag}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
CTCLIP_BENCH_STAGE_TABLE=gpurun_out/${tag}_stages.md CTCLIP_BENCH_GEMM_TABLE=gpurun_out/${tag}_gemm_table.txt \
  timeout 700 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> /dev/null; echo "ref rc=$?"
timeout 300 python bench.py --config zero_shot > gpurun_out/${tag}_bench_zero_shot.json 2> gpurun_out/${tag}_bench_zero_shot.err; echo "zero-shot rc=$?"
cut -c1-300 gpurun_out/${tag}_bench_zero_shot.json
timeout 300 python tools/attn_tc_probe.py --reps 10 > gpurun_out/${tag}_attn_probe.txt 2>&1
timeout 300 python tools/peg_stream_probe.py > gpurun_out/${tag}_peg_probe.txt 2>&1
# --set full captures on a 1+1-layer step (same kernels and shapes, fewer launches); summaries only
CMD="python bench.py --steps 1 --warmup 1 --depth 1 --bert-layers 1 --no-cpu-baseline --no-stages"
: > gpurun_out/${tag}_ncu_full_top_kernels.txt
for spec in "gemm_geglu:gemm_tc_kernel:8:1" "attn_tc_fwd:attn_tc_fwd_kernel:0:1" "attn_tc_bwd:attn_tc_bwd_kernel:0:1" "attn_dtab:attn_dtab_reduce:0:1" \
            "peg_conv:peg_stream_conv_kernel:0:4" "peg_wgrad:peg_stream_wgrad_kernel:0:2" "ln_bwd:ln_bwd_kernel:2:1" "geglu_bwd:geglu_bwd_kernel:0:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  timeout 300 ncu --set full --clock-control none -k "regex:$pat" -s $skip -c $cnt -f \
    -o /tmp/ncu_$tag/prof_${name} $CMD > /tmp/ncu_$tag/prof_${name}.log 2>&1
  echo "#### capture ${name} (ncu --set full --clock-control none -k regex:$pat -s $skip -c $cnt; $CMD)" >> gpurun_out/${tag}_ncu_full_top_kernels.txt
  python tools/ncu_summary.py /tmp/ncu_$tag/prof_${name}.ncu-rep >> gpurun_out/${tag}_ncu_full_top_kernels.txt 2>&1
done
grep -E "^####|^==|time  |dram_rd|dram_wr" gpurun_out/${tag}_ncu_full_top_kernels.txt | head -60
if [ "$2" != "quick" ]; then
  # launch list of one warm-up + one timed step (cold-cache serialised times: compare SHARES)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 2300 --log-file /tmp/ncu_$tag/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stages > /tmp/ncu_$tag/ncu_launch.log 2>&1; echo "launch list rc=$?"
  python tools/summarize_ncu_launches.py /tmp/ncu_$tag/launches.csv > gpurun_out/${tag}_launches.txt 2>&1; head -20 gpurun_out/${tag}_launches.txt
fi
du -sh gpurun_out
