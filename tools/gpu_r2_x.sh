mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "row_classes or column_class" 2>&1 | tail -2
for k in 1 2 4 8; do timeout 600 python tools/ab_mxv.py --scale 24 --shard $k --cfg "" --cfg "mxv_row_min_nnz=0" 2>&1 | cut -c1-250; done
