mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_fixup_merge=0" --cfg "mxv_fixup_merge=1" 2>&1 | grep "fixup\|cfg" | cut -c1-120
timeout 600 python tools/ab_mxv.py --scale 24 --shard 8 --profile --cfg "mxv_fixup_merge=0" --cfg "mxv_fixup_merge=1" --cfg "mxv_fixup_merge=1,mxv_phases=2" --cfg "mxv_fixup_merge=1,mxv_phases=3" --cfg "mxv_fixup_merge=1,mxv_row_classes=0" 2>&1 | cut -c1-220
