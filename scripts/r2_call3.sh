#!/usr/bin/env bash
# round 2, call 3: tcgen05 prepare (hang-guarded, first), vectorised prep/post parity, scan variants + timeline, bench
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scale_parity.py -q -x -k "tcgen05_prepare" --timeout 120 --timeout-method=thread > gpurun_out/t_prep.log 2>&1; tail -25 gpurun_out/t_prep.log
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_scale_parity.py tests/test_gpu_model.py tests/test_gpu_pretrained.py -q -k "gdn or model or checkpoint" --timeout 300 --timeout-method=thread > gpurun_out/t_gdn3.log 2>&1; tail -12 gpurun_out/t_gdn3.log
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune.log 2>&1; cat gpurun_out/scan_tune.log
KB2_GDN_PREPARE_MMA_SYNC=1 timeout 300 python bench.py --layers 8 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/b8_legacyprep.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mma.sync prepare:', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items() if k.startswith('gdn')})"
timeout 300 python bench.py --layers 8 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/b8_tcprep.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tcgen05 prepare:', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items() if k.startswith('gdn')})"
KB2_GDN_SCAN_SPLIT=1 timeout 300 python bench.py --layers 8 --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/b8_split.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split scan:', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kernel'].items() if k.startswith('gdn')})"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2c.json').read()); print(d['ms_per_step'], d['value']); print({k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel'].items()})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gdn_prepare_tc -s 3 -c 1 -f -o gpurun_out/prof_gdn_prepare_tc python bench.py --layers 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gdn_prep.log 2>&1; tail -2 gpurun_out/ncu_gdn_prep.log
