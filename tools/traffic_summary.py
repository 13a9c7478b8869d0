"""Development tool: per-kernel DRAM traffic and time of the LAST pipeline step in an ncu --csv metrics log.

    python tools/traffic_summary.py log.csv <launches per step>
"""
import csv
import json
import re
import sys
from collections import OrderedDict

path, per_step = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(line for line in open(path) if line.startswith('"'))]
hdr = rows[0]
ID, NAME, METRIC, VALUE = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
launches = OrderedDict()
for r in rows[1:]:
    launches.setdefault(int(r[ID]), {'name': r[NAME]})[r[METRIC]] = float(r[VALUE].replace(',', ''))
ids = sorted(launches)[-per_step:]
out = OrderedDict()
for i in ids:
    rec = launches[i]
    name = re.sub(r'[<(].*', '', rec['name'].replace('void ', '').replace('pml::', ''))
    k = out.setdefault(name, {'launches': 0, 'dram_bytes_read': 0., 'dram_bytes_write': 0., 'dram_bytes_total': 0., 'time_us_under_ncu': 0.})
    k['launches'] += 1
    k['dram_bytes_read'] += rec.get('dram__bytes_read.sum', 0.)
    k['dram_bytes_write'] += rec.get('dram__bytes_write.sum', 0.)
    k['dram_bytes_total'] = k['dram_bytes_read'] + k['dram_bytes_write']
    k['time_us_under_ncu'] += rec.get('gpu__time_duration.sum', 0.) / 1e3
tot = sum(k['time_us_under_ncu'] for k in out.values())
for k in out.values():
    k['share_of_step_time'] = k['time_us_under_ncu'] / tot
print(json.dumps({'what': 'DRAM traffic and time of ONE pipeline step (bottom-up + top-down sweeps) summed over all launches of '
                          'each kernel; ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum '
                          '--clock-control none, CUDA graphs disabled (tools/one_step.py)', 'kernels': out}, indent=1))
