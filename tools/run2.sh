set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/exp_mxv.py > gpurun_out/exp_mxv.txt 2>&1; cat gpurun_out/exp_mxv.txt
