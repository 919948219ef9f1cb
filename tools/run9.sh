mkdir -p gpurun_out
for d in 1.0 0.5; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_seg_$d.csv python tools/prof_phase.py --phases 4 --masked 1 --density $d > gpurun_out/ncu_l.log 2>&1
done
