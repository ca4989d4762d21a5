#!/usr/bin/env bash
# round 2, call 21 (2 GPUs): EP / communicator tests, GDN prep load-ahead check, N=2 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ep.py tests/test_gpu_attention.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_16.log 2>&1; tail -4 gpurun_out/t_16.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_qcn_n2_r02n.json 2> gpurun_out/bench_qcn_n2_r02n.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n2_r02n.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline'].get('component_ms_per_step'))"; tail -3 gpurun_out/bench_qcn_n2_r02n.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n1_r02n.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02n.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items() if 'prep' in k})"
