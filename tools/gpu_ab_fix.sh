mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vector_ops.py tests/test_gpu_parity.py -m gpu -x -q -k "vector_ops or pull or v_reduce or v_assign or v_eadd or v_count" 2>&1 | tail -3
timeout 200 python tools/ab_mxv.py --scale 24 --select ALWAYS --profile --cfg "" --out gpurun_out/ab_always24.jsonl 2>&1 | tail -12
timeout 200 python tools/ab_mxv.py --scale 22 --select ALWAYS --cfg "" --out gpurun_out/ab_always22.jsonl 2>&1 | tail -1
timeout 200 python tools/ab_mxv.py --scale 22 --cfg "" --out gpurun_out/ab_nqzero22.jsonl 2>&1 | tail -1
