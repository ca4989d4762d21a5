#!/usr/bin/env bash
# round 2, call 4: MMA issue microbenchmark, scan layout 2 (hang-guarded) + timelines, prepare_tc timeline, N=2 overlap check
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I krasis_b200/csrc scripts/ubench/mma_issue.cu -o /tmp/mma_issue && timeout 120 /tmp/mma_issue > gpurun_out/mma_issue.log 2>&1; cat gpurun_out/mma_issue.log
timeout 300 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05_scan" --timeout 120 --timeout-method=thread > gpurun_out/t_scan2.log 2>&1; tail -15 gpurun_out/t_scan2.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune2.log 2>&1; cat gpurun_out/scan_tune2.log
KB2_GDN_SCAN_LAYOUT=2 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_scale_parity.py -q -k "gdn" --timeout 300 --timeout-method=thread > gpurun_out/t_gdn4.log 2>&1; tail -6 gpurun_out/t_gdn4.log
