#!/usr/bin/env bash
# round 2, call 1: new tcgen05 GDN scan (hang-guarded) + the new scale-parity tests + a first bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05_scan" --timeout 120 --timeout-method=thread > gpurun_out/t_tc.log 2>&1; tail -15 gpurun_out/t_tc.log
timeout 600 python -m pytest tests/test_gpu_attention.py -q -k "gdn" --timeout 120 --timeout-method=thread > gpurun_out/t_gdn.log 2>&1; tail -8 gpurun_out/t_gdn.log
timeout 1200 python -m pytest tests/test_gpu_scale_parity.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_scale.log 2>&1; tail -25 gpurun_out/t_scale.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
