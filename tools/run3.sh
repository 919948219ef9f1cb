set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_v3.json 2> gpurun_out/bench_err.txt; tail -3 gpurun_out/bench_err.txt; cat gpurun_out/bench_r01_v3.json
timeout 900 ncu --kernel-name-base demangled -k regex:splacu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mxv_wtile_kernel -s 3 -c 1 -o gpurun_out/prof_mxv_wtile -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
