# gpurun -- 'bash tools/gpu_final.sh TAG': GPU parity suite, smoke(), both bench arms, the ncu launch list and one ncu --set full capture of a bench step
TAG=${1:-r17}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
timeout 600 python bench.py --extra > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'launches', j['gpu_launches'], 'cpu', j['cpu_baseline']['value'])
    print(json.dumps(j['extra'])[:1500])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_$TAG.err').read()[-3000:])
PY
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 14 -o gpurun_out/prof_step_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log | cut -c1-200
