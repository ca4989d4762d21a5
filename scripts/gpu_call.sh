timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_final_n1.json 2> gpurun_out/bench_r01_final_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_final_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
