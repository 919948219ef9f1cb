# round 2, call L: version 2 of the class kernel: parity of everything that runs through it, then A/B on RMAT-24
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "column_class or row_classes or tail or density or golden or known or baseline or config" 2>&1 | tail -5
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2l.jsonl \
  --cfg "mxv_seg_version=1" \
  --cfg "mxv_seg_version=2" \
  --cfg "mxv_seg_version=1,mxv_red=1" \
  --cfg "mxv_seg_version=2,mxv_red=1" \
  --cfg "mxv_seg_version=2,mxv_red=1,mxv_row_classes=2" \
  2>&1 | tail -8
timeout 600 python tools/ab_mxv.py --scale 24 --select ALWAYS --cfg "mxv_seg_version=1" --cfg "mxv_seg_version=2" --cfg "mxv_seg_version=2,mxv_red=1" 2>&1 | tail -4
export SPLACU_OPTIONS="mxv_seg_version=2"
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_mxv_r2l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/launches_mxv_r2l.csv') if l.startswith('"')))
for r in rows[-16:]:
    print(r['Kernel Name'][:70], r['Grid Size'], float(r['Metric Value'])/1000)
PY
