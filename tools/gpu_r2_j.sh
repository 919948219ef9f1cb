# round 2, call J: fused hub + tail kernel: parity, then A/B on RMAT-24 in one process
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -5
timeout 1200 python tools/ab_mxv.py --scale 24 --out gpurun_out/ab_r2j.jsonl \
  --cfg "" \
  --cfg "mxv_red=1" \
  --cfg "mxv_phase_slots=22528,mxv_phases=8" \
  --cfg "mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=8" \
  --cfg "mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=11" \
  --cfg "mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=10,mxv_fuse_warps=20" \
  --cfg "mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=8,mxv_red=1" \
  --cfg "mxv_phase_slots=16384,mxv_phases=11,mxv_fuse=8,mxv_red=1" \
  --cfg "mxv_phase_slots=28672,mxv_phases=6,mxv_fuse=8,mxv_red=1,mxv_fuse_smem_kb=160" \
  2>&1 | tail -12
