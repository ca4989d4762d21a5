#!/usr/bin/env bash
# round 2, call 20: prepare v2 with per-output bulk groups; bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_attention.py -q -x -m gpu -k "tcgen05 or gdn or GDN or delta" --timeout 300 --timeout-method=thread > gpurun_out/t_15.log 2>&1; tail -3 gpurun_out/t_15.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune13.log 2>&1; grep -v "^    \[" gpurun_out/scan_tune13.log | cut -c1-220; grep -A3 "v2 prepare" gpurun_out/scan_tune13.log | tail -2 | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02m.json 2> gpurun_out/bench_qcn_n1_r02m.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02m.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()})"
