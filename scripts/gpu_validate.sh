#!/usr/bin/env bash
# Round-end validation recipe on one B200 (run through `gpurun -- bash scripts/gpu_validate.sh`); everything it writes
# goes to gpurun_out/ (scratch) — copy what should be kept into profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of a reduced-depth run of the same command, then one full capture of the roofline kernel in the 48-layer step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full_8layers.csv \
    python bench.py --layers 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 60 -c 2 -f -o gpurun_out/prof_gemm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm_full.log 2>&1
