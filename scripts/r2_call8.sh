#!/usr/bin/env bash
# round 2, call 9: packed bf16 converts (no F2F) in GDN kernels; grouped GEMM per-item timeline
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05 or gdn" --timeout 120 --timeout-method=thread > gpurun_out/t_tc8.log 2>&1; tail -5 gpurun_out/t_tc8.log
timeout 300 python -m pytest tests/test_gpu_gdn.py tests/test_gpu_gqa.py -q -x --timeout 120 --timeout-method=thread > gpurun_out/t_gdn8.log 2>&1; tail -5 gpurun_out/t_gdn8.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune6.log 2>&1; cat gpurun_out/scan_tune6.log
timeout 300 python scripts/gemm_trace.py > gpurun_out/gemm_trace1.log 2>&1; cat gpurun_out/gemm_trace1.log
