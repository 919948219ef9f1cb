# round 2, call N: 24-warp hub CTAs for version 2; v1/v2 per-class timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_class_phases_auto or row_classes" 2>&1 | tail -3
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2n.jsonl \
  --cfg "mxv_seg_version=2,mxv_seg_warps=20" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=24" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=24,mxv_red=1" \
  --cfg "mxv_seg_version=1,mxv_red=1" \
  2>&1 | tail -5
export SPLACU_OPTIONS="mxv_seg_version=2,mxv_seg_warps=24,mxv_red=1"
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_mxv_r2n.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/launches_mxv_r2n.csv') if l.startswith('"')))
for r in rows[-16:]:
    print(r['Kernel Name'][:70], r['Grid Size'], r['Block Size'], float(r['Metric Value'])/1000)
PY
