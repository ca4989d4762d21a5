#!/usr/bin/env bash
# round 2, call 25: fused W8A8 chain, gqa_prep rewrite: attention + model tests, bench for QCN and q235b
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py tests/test_gpu_pretrained.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_20.log 2>&1; tail -3 gpurun_out/t_20.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02q.json 2> gpurun_out/bench_qcn_n1_r02q.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02q.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks'], d['roofline']['traffic']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline'].get('component_ms_per_step'))"
timeout 600 python bench.py --config q235b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q235b_n1_r02q.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_q235b_n1_r02q.json').read().strip().splitlines()[-1]); print('q235b', d['ms_per_step'], d['value']); print({k:round(v['ms_per_step'],2) for k,v in list(d['roofline']['per_kernel'].items())[:8]})"
