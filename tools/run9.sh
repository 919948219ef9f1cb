mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vxm_|bitmap|scan|select_bits|emit|fill_kernel|count" --csv --log-file gpurun_out/launches_vxm.csv python tools/prof_vxm.py 2>&1 | grep -v "^==" | tail -4
