timeout 240 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -x -q --timeout 60 --timeout-method=thread 2>&1 | grep -v "^  warn\|Warning" | tail -6
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full10.json 2> gpurun_out/bench_full10.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full10.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['component_ms_per_step'])
PY
