# round 2, call B: all GPU tests (no -x), bench with the structure-only push, ncu of the push kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "parity\]|passed|failed|rror|assert|^FAILED|^tests/" | cut -c1-400 | tail -60 ) 2>&1 | tail -64
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    print('vxm', json.dumps(j.get('vxm')))
    print('bfs', json.dumps(j.get('bfs')))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r2b.err').read()[-3000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. -s 200 -c 80 --csv --log-file gpurun_out/launches_vxm_r2b.csv python tools/prof_vxm.py > gpurun_out/prof_vxm_r2b.log 2>&1
tail -3 gpurun_out/prof_vxm_r2b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vxm_expand -s 4 -c 2 -o gpurun_out/prof_vxm_r2b -f python tools/prof_vxm.py > gpurun_out/ncu_vxm_r2b.log 2>&1
tail -1 gpurun_out/ncu_vxm_r2b.log | cut -c1-200
timeout 300 python tools/exp_bfs.py 2>&1 | cut -c1-200 | tail -24
