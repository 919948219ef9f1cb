# bank-ordered hub classes: parity of the pull paths, then the A/B with per-launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_bank_order=0" --cfg "mxv_bank_order=1" --out gpurun_out/ab_bank_order.jsonl 2>&1 | tail -40
