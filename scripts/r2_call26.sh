#!/usr/bin/env bash
# round 2, call 32: CUDA-graph replayed prefill step: parity test, N=1 bench with / without
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_scale_parity.py -q -x -m gpu -k "graph or tcgen05" --timeout 300 --timeout-method=thread > gpurun_out/t_26.log 2>&1; tail -12 gpurun_out/t_26.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qcn_n1_graph.json 2> gpurun_out/bench_qcn_n1_graph.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_graph.json').read().strip().splitlines()[-1]); print('graph', d['cuda_graph'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks']); print({k:round(v['ms_per_step'],2) for k,v in list(d['roofline']['per_kernel'].items())[:8]})" || tail -8 gpurun_out/bench_qcn_n1_graph.err
timeout 600 python bench.py --config qwen35 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qwen35_n1_graph.json 2> gpurun_out/bench_qwen35_n1_graph.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qwen35_n1_graph.json').read().strip().splitlines()[-1]); print('qwen35 graph', d['cuda_graph'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])" || tail -8 gpurun_out/bench_qwen35_n1_graph.err
