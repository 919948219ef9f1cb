# round 2, call M: ncu --set full of the class-0 and tail launches of version 2 (and version 1 of the tail for comparison)
mkdir -p gpurun_out
export SPLACU_OPTIONS="mxv_seg_version=2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mxv_seg2_kernel -s 15 -c 5 -o gpurun_out/prof_seg2_r2m -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/ncu_seg2_r2m.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_seg2_r2m.ncu-rep 2>/dev/null | grep -v "^  launch\|lts__t_sectors.sum\|dram__bytes_write"
