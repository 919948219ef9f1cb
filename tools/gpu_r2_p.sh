mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_class_phases or row_classes" 2>&1 | tail -3
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2p.jsonl \
  --cfg "mxv_seg_version=1" \
  --cfg "mxv_seg_version=2" \
  --cfg "mxv_seg_version=3" \
  --cfg "mxv_seg_version=1,mxv_red=1" \
  --cfg "mxv_seg_version=2,mxv_red=1" \
  --cfg "mxv_seg_version=3,mxv_red=1" \
  2>&1 | tail -8 | cut -c1-150
