"""tools/ncu_summary.py report.ncu-rep -- key metrics per profiled launch (reads the report with ncu here, no GPU)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'launch__registers_per_thread', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for d in data:
    print('----', d[idx['Kernel Name']][:110])
    for w in WANT:
        if w in idx:
            print(f"  {w:90s} {d[idx[w]]} {units[idx[w]]}")
    top = sorted(((float(d[idx[h]] or 0), h) for h in stalls), reverse=True)[:6]
    for val, h in top:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {val:.2f}")
