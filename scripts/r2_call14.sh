#!/usr/bin/env bash
# round 2, call 19: pruned gdn_tc.cu (final scan kernel only), prepare v2 with staged outputs + bulk stores
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_attention.py tests/test_gpu_model.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_14.log 2>&1; tail -4 gpurun_out/t_14.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune12.log 2>&1; cat gpurun_out/scan_tune12.log | cut -c1-250 | grep -v "^    \[" ; grep -A3 "v2 prepare" gpurun_out/scan_tune12.log | tail -2 | cut -c1-200
