set -x
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread 2>&1 | grep -v "^  warn\|Warning" | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_final_n1.json 2> gpurun_out/bench_r01_final_n1.err; tail -c 300 gpurun_out/bench_r01_final_n1.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; tail -c 200 gpurun_out/bench_r01_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full_8layers.csv python bench.py --layers 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 60 -c 2 -f -o gpurun_out/prof_gemm_full2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm_full2.log 2>&1
ls gpurun_out | tail -5
