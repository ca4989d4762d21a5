#!/usr/bin/env bash
# round 2, call 27 (2 GPUs): attention out_proj GEMM -> reduce-scatter fused over peer memory; exact division-free quantisation
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_zz_reference_fixtures.py tests/test_gpu_ep.py -q -x -m gpu --timeout 200 --timeout-method=thread > gpurun_out/t_22.log 2>&1; tail -6 gpurun_out/t_22.log | cut -c1-300
for fused in 1 0; do
KB2_FUSED_ATTN_RS=$fused timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$fused bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n2_attnrs$fused.json 2> gpurun_out/bench_qcn_n2_attnrs$fused.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n2_attnrs$fused.json').read().strip().splitlines()[-1]); print('fused_attn_rs=$fused', d['ms_per_step'], d['value']); print(d['roofline'].get('component_ms_per_step'))" || tail -8 gpurun_out/bench_qcn_n2_attnrs$fused.err
done
