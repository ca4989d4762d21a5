timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -x -q --timeout 90 --timeout-method=thread 2>&1 | grep -v "^  warn\|Warning" | tail -12
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full5.json 2> gpurun_out/bench_full5.err; tail -c 1200 gpurun_out/bench_full5.json
