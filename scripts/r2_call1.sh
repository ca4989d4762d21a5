#!/usr/bin/env bash
# round 2, call 1: new tcgen05 GDN scan (hang-guarded) + the new scale-parity tests + model tests + a first bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05_scan" --timeout 120 --timeout-method=thread > gpurun_out/t_tc.log 2>&1; tail -15 gpurun_out/t_tc.log
timeout 600 python -m pytest tests/test_gpu_attention.py -q -k "gdn" --timeout 120 --timeout-method=thread > gpurun_out/t_gdn.log 2>&1; tail -8 gpurun_out/t_gdn.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_pretrained.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_model.log 2>&1; tail -25 gpurun_out/t_model.log
timeout 1200 python -m pytest tests/test_gpu_scale_parity.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_scale.log 2>&1; tail -25 gpurun_out/t_scale.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json
timeout 600 python bench.py --config v2lite --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_v2lite.json 2> gpurun_out/bench_r2a_v2lite.err; tail -3 gpurun_out/bench_r2a_v2lite.err; cat gpurun_out/bench_r2a_v2lite.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gdn_scan_tc -s 3 -c 1 -f -o gpurun_out/prof_gdn_tc python bench.py --layers 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gdn_tc.log 2>&1; tail -3 gpurun_out/ncu_gdn_tc.log
