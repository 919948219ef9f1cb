mkdir -p gpurun_out
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_tail_hints=0" --cfg "mxv_tail_hints=1" --cfg "mxv_tail_hints=2" --cfg "mxv_tail_hints=3" 2>&1 | grep "class04\|row_class\|cfg" | cut -c1-120
