set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mxv_wtile_kernel -s 2 -c 1 -o gpurun_out/prof_mxv_wtile -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
