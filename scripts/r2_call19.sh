#!/usr/bin/env bash
# round 2, call 24 (4 GPUs): N=4 bench; gqa_prep rewrite check on the GQA tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_19.log 2>&1; tail -3 gpurun_out/t_19.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n4_r02p.json 2> gpurun_out/bench_qcn_n4_r02p.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n4_r02p.json').read().strip().splitlines()[-1]); print('N=4', d['ms_per_step'], d['value']); print(d['roofline'].get('component_ms_per_step')); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()})" || tail -5 gpurun_out/bench_qcn_n4_r02p.err
