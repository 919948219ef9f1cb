mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mxv_wtile" -c 4 -o gpurun_out/prof_phases2 -f python tools/prof_phase.py --phases 1 --masked 2 > gpurun_out/ncu_phases2.log 2>&1
tail -3 gpurun_out/ncu_phases2.log
