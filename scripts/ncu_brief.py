"""One line per kernel of an .ncu-rep: time, DRAM bytes/throughput, L1TEX/shared pipe, issue, occupancy."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
def col(r, name):
    try:
        return float(r[hdr.index(name)])
    except Exception:
        return float('nan')
for r in data:
    name = r[hdr.index('Kernel Name')][:28]
    t = col(r, 'gpu__time_duration.sum')
    rd, wr = col(r, 'dram__bytes_read.sum'), col(r, 'dram__bytes_write.sum')
    print(f"{name:28s} t={t:8.3f} rd={rd:8.3f} wr={wr:7.3f} "
          f"dram%={col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"l1tex%={col(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"smem%={col(r,'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):5.1f} "
          f"lsu%={col(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):5.1f} "
          f"l2hit={col(r,'lts__t_sector_hit_rate.pct'):5.1f} issue%={col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"warps%={col(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} regs={r[hdr.index('launch__registers_per_thread')]}")
