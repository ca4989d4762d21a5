#!/usr/bin/env bash
# round 2, call 31: refreshed bench lines for the four BASELINE configurations
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_qcn_n1_r02u.json 2> gpurun_out/bench_qcn_n1_r02u.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_qcn_n1_r02u.json').read().strip().splitlines()[-1]); print('qcn', d['ms_per_step'], d['value'], d['clocks'], d['gpu_launches']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline'].get('component_ms_per_step'))"
for cfg in v2lite qwen35 q235b; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_n1_r02u.json 2> gpurun_out/bench_${cfg}_n1_r02u.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_${cfg}_n1_r02u.json').read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['value'], d['clocks']['sm_mhz']); print({k:round(v['ms_per_step'],2) for k,v in list(d['roofline']['per_kernel'].items())[:8]})" || tail -3 gpurun_out/bench_${cfg}_n1_r02u.err
done
