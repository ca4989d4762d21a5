#!/usr/bin/env bash
# round 2, call 30: prepare v2 with the inverse's block products on mma.sync 3xTF32
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_attention.py -q -x -m gpu -k "tcgen05 or gdn or GDN or delta" --timeout 300 --timeout-method=thread > gpurun_out/t_24.log 2>&1; tail -3 gpurun_out/t_24.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune15.log 2>&1; grep -E "^GDN|v2 prepare" gpurun_out/scan_tune15.log | cut -c1-250; grep -A4 "v2 prepare" gpurun_out/scan_tune15.log | tail -3 | cut -c1-200
