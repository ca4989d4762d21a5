#!/usr/bin/env bash
# round 2, call 33 (2 GPUs): graph-owned KV sequence fix; N=2 bench with the CUDA-graph replayed step vs eager launches
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ep.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_27.log 2>&1; tail -8 gpurun_out/t_27.log | cut -c1-300
for g in graph nograph; do
flag=""; [ "$g" = "nograph" ] && flag="--no-graph"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline $flag > gpurun_out/bench_qcn_n2_$g.json 2> gpurun_out/bench_qcn_n2_$g.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n2_$g.json').read().strip().splitlines()[-1]); print('$g', d['cuda_graph'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'])" || tail -12 gpurun_out/bench_qcn_n2_$g.err
done
