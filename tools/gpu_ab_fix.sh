mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pull or op_pairs or named" 2>&1 | tail -4
timeout 600 python tools/ab_mxv.py --scale 24 --profile --cfg "mxv_fixup_merge=1" --cfg "mxv_fixup_merge=2" --out gpurun_out/ab_fixup_flat.jsonl 2>&1 | grep -E "fixup|cfg"
timeout 300 tools/bin/stream_gather_bench 2>&1 | tee gpurun_out/stream_gather_bench.txt
