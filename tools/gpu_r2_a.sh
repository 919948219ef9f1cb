# round 2, call A: tightened parity tests, the new bench line (vxm + bfs blocks), the reference arm at scale 24, baseline ncu of the push kernel
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
free -g | head -2; nproc
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "parity\]|passed|failed|error|Error|assert" | tail -40 ) 2>&1 | tail -45
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'parity', j['parity'])
    print('vxm', json.dumps(j.get('vxm')))
    print('bfs', json.dumps(j.get('bfs')))
    print('cpu', json.dumps(j.get('cpu_baseline')))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r2a.err').read()[-3000:])
PY
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err ) 2>&1 | tail -3
tail -c 1500 gpurun_out/bench_ref_r2a.json; tail -5 gpurun_out/bench_ref_r2a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vxm_ -c 60 --csv --log-file gpurun_out/launches_vxm_r2a.csv python tools/prof_vxm.py > gpurun_out/prof_vxm_r2a.log 2>&1
tail -3 gpurun_out/prof_vxm_r2a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vxm_expand -s 2 -c 1 -o gpurun_out/prof_vxm_r2a -f python tools/prof_vxm.py > gpurun_out/ncu_vxm_r2a.log 2>&1
tail -1 gpurun_out/ncu_vxm_r2a.log | cut -c1-200
