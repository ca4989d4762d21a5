#!/usr/bin/env bash
# round 2, call 15: scan layout 5, GEMM 4 token stages + hoisted routing-weight loads
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_moe.py tests/test_gpu_loader.py -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_10.log 2>&1; tail -4 gpurun_out/t_10.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune8.log 2>&1; grep -E "^layout|period|prepare:|max" gpurun_out/scan_tune8.log; grep -A3 "layout=5" gpurun_out/scan_tune8.log | tail -2
timeout 300 python scripts/gemm_trace.py > gpurun_out/gemm_trace3.log 2>&1; head -6 gpurun_out/gemm_trace3.log; sed -n 16,22p gpurun_out/gemm_trace3.log
KB2_GDN_SCAN_LAYOUT=5 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02j.json 2> gpurun_out/bench_qcn_n1_r02j.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02j.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()})"
