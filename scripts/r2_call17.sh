#!/usr/bin/env bash
# round 2, call 22: smoke, other BASELINE configurations, reference arm, ncu launch list + full captures for kernel_traffic.json
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
for cfg in v2lite qwen35 q235b; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_n1_r02o.json 2> gpurun_out/bench_${cfg}_n1_r02o.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_${cfg}_n1_r02o.json').read().strip().splitlines()[-1]); print('$cfg', d['ms_per_step'], d['value'], d['config'].get('tokens')); print({k:round(v['ms_per_step'],2) for k,v in list(d['roofline']['per_kernel'].items())[:8]})" || tail -3 gpurun_out/bench_${cfg}_n1_r02o.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02o.json 2> gpurun_out/bench_ref_r02o.err; cat gpurun_out/bench_ref_r02o.json | cut -c1-600
timeout 400 python bench.py --impl reference --cpu-format gguf --steps 2 --warmup 1 > gpurun_out/bench_ref_gguf_r02o.json 2> gpurun_out/bench_ref_gguf_r02o.err; cat gpurun_out/bench_ref_gguf_r02o.json | cut -c1-600
# launch list of a reduced-depth run of the bench command, then full captures of the heaviest kernel classes inside the 48-layer step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_launches_8layers.csv \
    python bench.py --layers 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1; tail -2 gpurun_out/ncu_l8.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_gemm_kernel|grouped_gemm_kernel|gdn_scan_tc_kernel|gdn_prepare_tc2_kernel|gqa_fmha_kernel" -s 40 -c 14 -f -o gpurun_out/r02o_top_kernels \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/r02o_top_kernels.ncu-rep --page raw --csv > gpurun_out/r02o_top_kernels_raw.csv 2>/dev/null; wc -l gpurun_out/r02o_top_kernels_raw.csv
