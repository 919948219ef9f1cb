mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_final2.err').read()[-3000:])
PY
