mkdir -p gpurun_out
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2r.jsonl \
  --cfg "mxv_red=1" --cfg "mxv_red=0" --cfg "mxv_red=1,mxv_row_classes=2" --cfg "mxv_red=1,mxv_phases=5" --cfg "mxv_red=1,mxv_phases=3" 2>&1 | tail -6 | cut -c1-230
timeout 600 python tools/ab_mxv.py --scale 24 --select ALWAYS --cfg "mxv_red=1" --cfg "mxv_red=0" 2>&1 | tail -3 | cut -c1-150
timeout 600 python tools/ab_mxv.py --scale 22 --cfg "mxv_red=1" --cfg "mxv_red=0" 2>&1 | tail -3 | cut -c1-150
