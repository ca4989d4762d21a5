#!/usr/bin/env bash
# round 2, call 17: defaults = tcgen05 prepare v2 + scan layout 5; router column blocks; pipelined GEMM epilogue. Full suite + bench.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu --timeout 300 --timeout-method=thread > gpurun_out/t_12.log 2>&1; tail -4 gpurun_out/t_12.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune10.log 2>&1; grep -E "prepare:|unit period" gpurun_out/scan_tune10.log; grep -A3 "v2 prepare" gpurun_out/scan_tune10.log | tail -2
timeout 300 python scripts/gemm_trace.py > gpurun_out/gemm_trace4.log 2>&1; head -5 gpurun_out/gemm_trace4.log; sed -n 16,20p gpurun_out/gemm_trace4.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02k.json 2> gpurun_out/bench_qcn_n1_r02k.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()})"
