mkdir -p gpurun_out
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_r11.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 14 -o gpurun_out/prof_step_r11 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r11.log 2>&1
tail -2 gpurun_out/ncu_full_r11.log | cut -c1-200
