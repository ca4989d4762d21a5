#!/usr/bin/env bash
# Round-end validation recipe on one B200 (run through `gpurun -- bash scripts/gpu_validate.sh`); everything it writes
# goes to gpurun_out/ (scratch) — copy what should be kept into profiles/.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline'].get('component_ms_per_step'))"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
# launch list of a reduced-depth run of the same command, then full captures of the heaviest kernel classes inside the 48-layer step
# (profiles/ncu_traffic.py turns the raw page into profiles/kernel_traffic.json)
if [ "${KB2_VALIDATE_NCU:-0}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_8layers.csv \
    python bench.py --layers 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_gemm_kernel|grouped_gemm_kernel|gdn_scan_tc_kernel|gdn_prepare_tc_kernel|gqa_fmha_kernel" -s 40 -c 14 -f -o gpurun_out/top_kernels \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/top_kernels.ncu-rep --page raw --csv > gpurun_out/top_kernels_raw.csv 2>/dev/null
fi
