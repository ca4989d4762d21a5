#!/usr/bin/env bash
# round 2, call 14: warp-uniform MMA/TMA issue in every tcgen05 kernel: full GPU suite, timelines, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_all9.log 2>&1; tail -8 gpurun_out/t_all9.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune7.log 2>&1; grep -E "^layout|period|prepare:|max" gpurun_out/scan_tune7.log
timeout 300 python scripts/gemm_trace.py > gpurun_out/gemm_trace2.log 2>&1; head -8 gpurun_out/gemm_trace2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02i.json 2> gpurun_out/bench_qcn_n1_r02i.err; cat gpurun_out/bench_qcn_n1_r02i.json; tail -3 gpurun_out/bench_qcn_n1_r02i.err
