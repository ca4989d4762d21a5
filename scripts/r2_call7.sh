#!/usr/bin/env bash
# round 2, call 8: scan layout 4 (hi/lo stacked along N) A/B tests + timelines
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05" --timeout 120 --timeout-method=thread > gpurun_out/t_tc7.log 2>&1; tail -8 gpurun_out/t_tc7.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune5.log 2>&1; cat gpurun_out/scan_tune5.log
