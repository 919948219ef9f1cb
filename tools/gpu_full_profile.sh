# all GPU tests, the default bench line, the ncu launch list and a --set full capture of one whole step
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
( time timeout 420 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|rror|assert|^FAILED|^tests/" | cut -c1-300 | tail -20 ) 2>&1 | tail -24
timeout 100 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -1
timeout 420 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    print('vxm', json.dumps(j.get('vxm'))[:600])
    print('bfs', json.dumps(j.get('bfs'))[:300])
    print('plugin', json.dumps(j.get('plugin'))[:600])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_final.err').read()[-3000:])
PY
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_mxv_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/bench_under_ncu_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 48 -c 16 -o gpurun_out/prof_step_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/ncu_full_final.log 2>&1
tail -1 gpurun_out/ncu_full_final.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/prof_step_final.ncu-rep > gpurun_out/ncu_step_final_summary.txt 2>/dev/null
grep -E "^----|gpu__time_duration|dram__bytes" gpurun_out/ncu_step_final_summary.txt | cut -c1-150
