"""Top stalled SASS instructions of a kernel in an .ncu-rep: python scripts/ncu_hot.py rep [section] [n]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
i0 = secs[k]; i1 = secs[k + 1] if k + 1 < len(secs) else len(rows)
hdr = rows[i0]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
cols = [c for c in ('stall_long_sb', 'stall_lg', 'stall_short_sb', 'stall_mio', 'stall_wait', 'stall_barrier', 'stall_not_selected', 'stall_math') if c in hdr]
body = [r for r in rows[i0 + 1:i1] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in body)
print('kernel section', k, 'of', len(secs), 'total samples', tot)
tots = {c: sum(int(r[hdr.index(c)] or 0) for r in body) for c in cols}
print('stall totals', {c: round(100 * v / max(tot, 1), 1) for c, v in tots.items()})
for r in sorted(body, key=lambda r: -int(r[isamp]))[:n]:
    print(f"{100*int(r[isamp])/tot:5.1f}%  exec {r[iex]:>10}  {r[isrc][:100]}")
