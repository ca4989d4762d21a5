#!/usr/bin/env bash
# round 2, call 2 (2 GPUs): loaders / tile cache / engine hand-off / new GGUF block types, EP tests on 2 GPUs, N=2 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_moe.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_load.log 2>&1; tail -15 gpurun_out/t_load.log
timeout 600 python -m pytest tests/test_gpu_ep.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_ep.log 2>&1; tail -15 gpurun_out/t_ep.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err; tail -5 gpurun_out/bench_r2b_n2.err; cat gpurun_out/bench_r2b_n2.json
timeout 600 python bench.py --config qwen35 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_qwen35.json 2> gpurun_out/bench_r2b_qwen35.err; tail -3 gpurun_out/bench_r2b_qwen35.err; cat gpurun_out/bench_r2b_qwen35.json | cut -c1-1500
timeout 900 python bench.py --config q235b --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_q235b.json 2> gpurun_out/bench_r2b_q235b.err; tail -3 gpurun_out/bench_r2b_q235b.err; cat gpurun_out/bench_r2b_q235b.json | cut -c1-1500
