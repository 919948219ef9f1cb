# round 2, call I: ncu --set full of the row-class kernel and of the tail class with row classes on
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mxv_scat_kernel -s 3 -c 1 -o gpurun_out/prof_scat_r2i -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/ncu_scat_r2i.log 2>&1
tail -2 gpurun_out/ncu_scat_r2i.log | cut -c1-200
ncu -i gpurun_out/prof_scat_r2i.ncu-rep --page raw --csv > gpurun_out/prof_scat_r2i_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_scat_r2i_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.sum','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','smsp__inst_executed_op_shared_atom.sum','smsp__inst_executed_op_shared_ld.sum','sm__cycles_elapsed.max','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for i,h in enumerate(hdr):
    if h in want or 'shared' in h and ('atom' in h or 'bank' in h):
        print(h, units[i], vals[i])
PY
