mkdir -p gpurun_out
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_red=1,mxv_phase_only=5" --cfg "mxv_red=1,mxv_phase_only=1" --cfg "mxv_red=1,mxv_phase_only=2"  --cfg "mxv_red=1,mxv_phase_only=6" --cfg "mxv_red=1,mxv_l2_persist=1" 2>&1 | grep -v "fixup\|hub_pack\|csr_pass" | cut -c1-120
