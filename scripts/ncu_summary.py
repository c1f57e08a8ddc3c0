"""Print the key metrics of every kernel in an .ncu-rep (uses `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
for r in rows[2:]:
    print('---')
    for w, i in idx:
        print(f'  {w:75s} {r[i]} {rows[1][i]}')
if len(sys.argv) > 2:
    for i, h in enumerate(hdr):
        if sys.argv[2] in h:
            print(h, [r[i] for r in rows[1:]])
