set -x
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_loader.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_moe.py -x -q -k "gguf" 2>&1 | tail -8
timeout 300 python profiles/gdn_probe.py 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gdn2.csv python profiles/gdn_probe.py 1 > gpurun_out/ncu_gdn2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdn_chunk -s 2 -c 2 -f -o gpurun_out/prof_gdn2 python profiles/gdn_probe.py 1 > gpurun_out/ncu_gdn2f.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full4.json 2> gpurun_out/bench_full4.err; tail -c 1500 gpurun_out/bench_full4.json
