#!/bin/bash
# GPU box: timing of the attention kernels + ncu captures of the tcgen05 forward / backward kernels + the new unit tests
TAG=${1:-r2b}
mkdir -p gpurun_out
python tools/attn_tc_probe.py --reps 10 > gpurun_out/${TAG}_attn_probe.txt 2>&1
cat gpurun_out/${TAG}_attn_probe.txt
ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 1 -c 1 -o gpurun_out/${TAG}_ncu_attn_fwd -f \
    python tools/attn_tc_probe.py --reps 1 --only tc --fwd-only > gpurun_out/${TAG}_ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -s 1 -c 1 -o gpurun_out/${TAG}_ncu_attn_bwd -f \
    python tools/attn_tc_probe.py --reps 1 --only tc > gpurun_out/${TAG}_ncu_bwd.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_fwd.log gpurun_out/${TAG}_ncu_bwd.log
python -m pytest tests/test_attention_tc_gpu.py tests/test_bert_gpu.py tests/test_gemm_gpu.py tests/test_headline_geometry_gpu.py tests/test_preprocess_gpu.py tests/test_retrieval_gpu.py tests/test_kernels_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_new.txt
tail -25 gpurun_out/${TAG}_pytest_new.txt
