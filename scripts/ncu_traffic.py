"""Aggregate DRAM traffic and time per kernel from an `ncu --set full` capture of ONE batch:
python scripts/ncu_traffic.py capture.ncu-rep out.json   (feeds bench.py's roofline.traffic)"""
import csv, json, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
def val(r, name, unit_scale):
    i = hdr.index(name)
    return float(r[i]) * unit_scale.get(units[i], 1.0)
BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
TIME = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
res = {}
for r in data:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').split('<')[0]
    ms = val(r, 'gpu__time_duration.sum', TIME)
    b = val(r, 'dram__bytes_read.sum', BYTES) + val(r, 'dram__bytes_write.sum', BYTES)
    k = res.setdefault(name, {'launches': 0, 'launches_nonempty': 0, 'ms': 0.0, 'dram_bytes': 0.0, 'per_launch_ms_GB': []})
    k['launches'] += 1
    k['launches_nonempty'] += 1 if b > 1e6 else 0
    k['ms'] += ms
    k['dram_bytes'] += b
    k['per_launch_ms_GB'].append([round(ms, 3), round(b / 1e9, 3)])
json.dump(res, open(sys.argv[2], 'w'), indent=1)
for k, v in res.items():
    print(k, v['launches'], 'launches', round(v['ms'], 2), 'ms', round(v['dram_bytes'] / 1e9, 1), 'GB')
