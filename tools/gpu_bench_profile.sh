mkdir -p gpurun_out
timeout 600 python bench.py --extra > gpurun_out/bench_r16.json 2> gpurun_out/bench_r16.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r16.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'launches', j['gpu_launches'])
    print(json.dumps(j['extra']))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r16.err').read()[-3000:])
PY
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_r16.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 14 -o gpurun_out/prof_step_r16 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r16.log 2>&1
tail -1 gpurun_out/ncu_full_r16.log | cut -c1-200
timeout 600 python tools/exp_bfs.py 2>&1 | cut -c1-200 | tail -24
