mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_fixup_merge=1" --cfg "mxv_fixup_merge=2" --out gpurun_out/ab_fixup_rows.jsonl 2>&1 | grep -E "fixup|cfg"
timeout 900 python bench.py --no-plugin --no-bfs > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -c 600 gpurun_out/bench_q.err; python - <<PY
import json
j=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
PY
