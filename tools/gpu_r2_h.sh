# round 2, call H: row classes of the tail (mxv_scat.cu): parity, A/B against no row classes, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "row_classes or ingest or column_class or tail or density or edge or golden or known" 2>&1 | tail -8
for opt in "mxv_row_classes=0" "mxv_row_classes=1" "mxv_row_classes=2" "mxv_row_classes=1,mxv_red=1"; do
  SPLACU_OPTIONS=$opt timeout 600 python bench.py --no-cpu-baseline --no-vxm --no-bfs --no-plugin --steps 30 > gpurun_out/bench_r2h_$opt.json 2> gpurun_out/bench_r2h_$opt.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r2h_$opt.json').read().strip().splitlines()[-1])
    print('$opt', 'value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'parity', j['parity']['rel_diff'])
    print(j['config']['kernel'])
except Exception as e:
    print('$opt', 'fail', e); print(open('gpurun_out/bench_r2h_$opt.err').read()[-2000:])
PY
done
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
SPLACU_OPTIONS=mxv_row_classes=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_mxv_r2h.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/launches_mxv_r2h.csv') if l.startswith('"')))
for r in rows[-22:]:
    print(r['Kernel Name'][:60], r['Grid Size'], float(r['Metric Value'])/1000)
PY
