#!/usr/bin/env bash
# round 2, call 5: double-buffered down GEMM + gather4 (hang-guarded), integer bf16 split in the scan, bench variants
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moe.py -q -x --timeout 120 --timeout-method=thread > gpurun_out/t_moe5.log 2>&1; tail -15 gpurun_out/t_moe5.log
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_model.py tests/test_gpu_loader.py -q --timeout 300 --timeout-method=thread > gpurun_out/t_scale5.log 2>&1; tail -8 gpurun_out/t_scale5.log
timeout 300 python scripts/gdn_scan_tune.py 2>&1 | grep -v "^    \[" > gpurun_out/scan_tune3.log; cat gpurun_out/scan_tune3.log
for v in "base:" "gather:KB2_MOE_GATHER=1" "layout2:KB2_GDN_SCAN_LAYOUT=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --layers 8 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/b8_$name.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items() if k.startswith('grouped') or k.startswith('binning') or k.startswith('gdn_chunk') or k.startswith('combine')})"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -3 gpurun_out/bench_r2e.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2e.json').read()); print(d['ms_per_step'], d['value']); print({k:(round(v['ms_per_step'],2), round(v.get('frac_of_peak',0),2)) for k,v in d['roofline']['per_kernel'].items()})"
