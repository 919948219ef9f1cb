set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 ./tools/gather_bench > gpurun_out/gather_bench.txt 2>&1; tail -40 gpurun_out/gather_bench.txt
timeout 600 python tools/exp_mxv.py > gpurun_out/exp_mxv.txt 2>&1; cat gpurun_out/exp_mxv.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_v2.json 2> gpurun_out/bench_err.txt; tail -3 gpurun_out/bench_err.txt; cat gpurun_out/bench_r01_v2.json
