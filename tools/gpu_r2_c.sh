# round 2, call C: all GPU tests (user ops via NVRTC, structure-only push with the coarse index), bench, ncu of the push kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|rror|assert|^FAILED|^tests/" | cut -c1-500 | tail -40 ) 2>&1 | tail -44
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    print('vxm', json.dumps(j.get('vxm')))
    print('bfs', json.dumps(j.get('bfs')))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r2c.err').read()[-3000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vxm_|bitmap_|scan_|select_bits' -c 80 --csv --log-file gpurun_out/launches_vxm_r2c.csv python tools/prof_vxm.py > gpurun_out/prof_vxm_r2c.log 2>&1
tail -3 gpurun_out/prof_vxm_r2c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vxm_expand_struct -s 2 -c 1 -o gpurun_out/prof_vxm_r2c -f python tools/prof_vxm.py > gpurun_out/ncu_vxm_r2c.log 2>&1
tail -1 gpurun_out/ncu_vxm_r2c.log | cut -c1-200
timeout 300 python tools/exp_bfs.py 2>&1 | cut -c1-200 | tail -24
