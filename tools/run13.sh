mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r13.json 2> gpurun_out/bench_r13.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r13.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e'], 'cpu', j['cpu_baseline'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r13.err').read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-800
