#!/usr/bin/env bash
# round 2, call 18: scan layout 6 (transposed v tile write) A/B + timeline
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale_parity.py -q -x -m gpu -k "tcgen05" --timeout 300 --timeout-method=thread > gpurun_out/t_13.log 2>&1; tail -4 gpurun_out/t_13.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune11.log 2>&1; grep -E "^layout|period|max" gpurun_out/scan_tune11.log; grep -A3 "layout=6" gpurun_out/scan_tune11.log | tail -2 | cut -c1-200
