timeout 600 ncu --set full --clock-control none --import-source on -k regex:vxm_expand_kernel -s 1 -c 1 -o gpurun_out/prof_vxm_expand -f python tools/exp_bfs.py > gpurun_out/ncu_vxm.log 2>&1
tail -3 gpurun_out/ncu_vxm.log
