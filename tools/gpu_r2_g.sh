# round 2, call G: A/B of the L2-reduction hand-over (mxv_red) and parity of it
mkdir -p gpurun_out
for opt in "mxv_red=0" "mxv_red=1"; do
  SPLACU_OPTIONS=$opt timeout 600 python bench.py --no-cpu-baseline --no-vxm --no-bfs --no-plugin --steps 30 > gpurun_out/bench_r2g_$opt.json 2> gpurun_out/bench_r2g_$opt.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_r2g_$opt.json').read().strip().splitlines()[-1])
print('$opt', 'value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'parity', j['parity']['rel_diff'])
PY
done
SPLACU_OPTIONS=mxv_red=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
SPLACU_OPTIONS=mxv_red=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_mxv_r2g_red.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > /dev/null 2>&1
