set -x
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_moe.py -x -q -k "gguf" 2>&1 | grep -E "Error|assert|beyond|passed|failed" | head -20
timeout 300 python profiles/gdn_probe.py 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gdn3.csv python profiles/gdn_probe.py 1 > gpurun_out/ncu_gdn3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gdn_chunk_scan -s 1 -c 1 -f -o gpurun_out/prof_gdn3 python profiles/gdn_probe.py 1 > gpurun_out/ncu_gdn3f.log 2>&1
