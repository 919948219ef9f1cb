mkdir -p gpurun_out
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_red=1" --cfg "mxv_red=0" 2>&1 | tail -40 | cut -c1-200
