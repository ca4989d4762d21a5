#!/usr/bin/env bash
# round 2, call 16: prepare v2 (eight core warps, no exchange) A/B + timelines; scan layout 5 fine stamps
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_attention.py -q -x -m gpu -k "tcgen05 or gdn or GDN or delta" --timeout 300 --timeout-method=thread > gpurun_out/t_11.log 2>&1; tail -4 gpurun_out/t_11.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune9.log 2>&1; cat gpurun_out/scan_tune9.log | cut -c1-330
