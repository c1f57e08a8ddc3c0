"""Instruction mix of one kernel launch in an .ncu-rep, by SASS opcode, normalised per gathered row
(= executed LDGSTS): python scripts/ncu_opmix.py rep kernel-regex launch-index"""
import csv, re, subprocess, sys, collections
rep, rx, idx = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f'::regex:{rx}:{idx}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[i0]
isrc, iex, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix, samp = collections.Counter(), collections.Counter()
for r in rows[i0 + 1:]:
    if r and r[0] == 'Address':
        break  # the page repeats the listing: count it once
    if len(r) <= iex or not r[iex].isdigit():
        continue
    m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', r[isrc])
    op = m.group(1).split('.')[0] if m else '?'
    if op in ('LDG', 'STG', 'LDS', 'STS', 'ATOMS', 'ATOMG', 'RED', 'SHFL'):
        op = m.group(1) if op in ('LDG', 'LDS') else op
    mix[op] += int(r[iex]); samp[op] += int(r[isamp] or 0)
total = sum(mix.values()); rows_g = mix.get('LDGSTS', 0) or 1
tsamp = sum(samp.values()) or 1
print(f"{rows[0][1][:70]}  launch {idx}: {total} warp instructions, {rows_g} LDGSTS (gathered rows); {total/rows_g:.1f} per row")
for op, c in mix.most_common(28):
    print(f"  {op:22s} {c:12d}  {c/rows_g:6.2f}/row  {100*c/total:5.1f}% of issued  {100*samp[op]/tsamp:5.1f}% of stall samples")
