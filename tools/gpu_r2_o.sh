# round 2, call O: version 2 on the hub classes only
mkdir -p gpurun_out
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2o.jsonl \
  --cfg "mxv_seg_version=1" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=20" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=24" \
  --cfg "mxv_seg_version=1,mxv_red=1" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=20,mxv_red=1" \
  --cfg "mxv_seg_version=2,mxv_seg_warps=24,mxv_red=1" \
  --cfg "mxv_seg_version=3,mxv_seg_warps=24,mxv_red=1" \
  2>&1 | tail -8
