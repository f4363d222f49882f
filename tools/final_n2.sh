#!/bin/bash
# 2-GPU box: inference tests + zero-shot bench (prefetching host pipeline), then bench.py --gpus 2 under different NCCL CTA caps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py -q 2>&1 | tail -3
timeout 400 python bench.py --config zero_shot > gpurun_out/r2r_bench_zero_shot.json 2> gpurun_out/r2r_bench_zero_shot.err; echo "zero-shot rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2r_bench_zero_shot.json').read().strip().splitlines()[-1]);print('zero-shot value',d['value'],'e2e',d['e2e']['value'])"
for ctas in default 4 8 16; do
  if [ "$ctas" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$((RANDOM%10)) bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r2r_bench_n2_ctas_${ctas}.json 2> gpurun_out/r2r_bench_n2_ctas_${ctas}.err
  python -c "import json,sys;d=json.loads(open('gpurun_out/r2r_bench_n2_ctas_${ctas}.json').read().strip().splitlines()[-1]);print('NCCL_MAX_CTAS=${ctas}: ms/step',d['ms_per_step'],'volumes/s',d['value'],'clocks',d['clocks']['sm_mhz'])"
done
