mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_algorithms.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "" --out gpurun_out/ab_side.jsonl 2>&1 | tail -14
timeout 600 python tools/ab_mxv.py --scale 24 --shard 8 --profile --cfg "" 2>&1 | tail -14
