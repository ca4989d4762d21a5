#!/usr/bin/env bash
# round 2, call 35 (4 GPUs): N=4 with the fused expert reduce-scatter and the graph-replayed step
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n4_r02w.json 2> gpurun_out/bench_qcn_n4_r02w.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n4_r02w.json').read().strip().splitlines()[-1]); print('N=4', d['cuda_graph'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step']); print(d['roofline'].get('component_ms_per_step'))" || tail -8 gpurun_out/bench_qcn_n4_r02w.err
