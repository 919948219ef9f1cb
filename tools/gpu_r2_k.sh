# round 2, call K: launch list + ncu full of the fused kernel
mkdir -p gpurun_out
export SPLACU_OPTIONS="mxv_phase_slots=22528,mxv_phases=8,mxv_fuse=8"
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/launches_mxv_r2k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(l for l in open('gpurun_out/launches_mxv_r2k.csv') if l.startswith('"')))
for r in rows[-24:]:
    print(r['Kernel Name'][:70], r['Grid Size'], float(r['Metric Value'])/1000)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mxv_seg_fused -s 3 -c 1 -o gpurun_out/prof_fused_r2k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/ncu_fused_r2k.log 2>&1

python tools/ncu_summary.py gpurun_out/prof_fused_r2k.ncu-rep 2>/dev/null | head -40
