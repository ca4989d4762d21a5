#!/usr/bin/env bash
# round 2, call 23 (2 GPUs): chunk-pipelined attention collectives: 2-rank model vs 1-rank model, N=2 bench with and without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ep.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_18.log 2>&1; tail -4 gpurun_out/t_18.log
for pipe in 1 0; do
KB2_PIPELINE_ATTENTION=$pipe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pipe bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n2_pipe$pipe.json 2> gpurun_out/bench_qcn_n2_pipe$pipe.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n2_pipe$pipe.json').read().strip().splitlines()[-1]); print('pipe=$pipe', d['ms_per_step'], d['value']); print(d['roofline'].get('component_ms_per_step'))" || tail -5 gpurun_out/bench_qcn_n2_pipe$pipe.err
done
