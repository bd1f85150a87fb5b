"""Per-kernel DRAM traffic and achieved bandwidth from an ncu CSV of one eager training step:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s N -c M --csv \
        --log-file gpurun_out/step_traffic.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table
    python tools/membound_summary.py gpurun_out/step_traffic.csv [steps] > profiles/r02_step_dram_traffic.txt

Times under ncu are serialised and cold-cache (each launch starts with whatever the previous one left in L2): GB/s here is a
LOWER bound on what the kernel reaches inside the replayed graph; the byte counts are exact.
"""
import collections
import csv
import json
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mn, mu, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value'), hdr.index('ID')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', '')) * scale.get(r[mu], 1.0)
    except ValueError:
        continue
    d = per.setdefault(r[idc], {'name': re.sub(r'^void ', '', re.sub(r'\(.*', '', r[kn]))[:90]})
    d[r[mn]] = v
# whole steps only: keep the launches after the first p2p_final_kernel (the end of a loss forward) up to and including the last one
marks = [i for i, d in enumerate(per.values()) if 'p2p_final_kernel' in d['name']]
launches = list(per.values())
if steps <= 0 and len(marks) >= 2:
    launches = launches[marks[0] + 1:marks[-1] + 1]
    steps = float(len(marks) - 1)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in launches:
    a = agg[d['name']]
    a[0] += 1
    a[1] += d.get('gpu__time_duration.sum', 0.0)
    a[2] += d.get('dram__bytes_read.sum', 0.0)
    a[3] += d.get('dram__bytes_write.sum', 0.0)
if steps <= 0:
    steps = 1.0
peak = 6546.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
tt = sum(a[1] for a in agg.values())
print('# %d launches (%g step[s]); device time %.1f us/step; DRAM read %.1f MB + write %.1f MB per step; HBM peak %.0f GB/s (MEASURED_PEAKS.json)' % (
    len(launches), steps, tt * 1e6 / steps, sum(a[2] for a in agg.values()) / 1e6 / steps, sum(a[3] for a in agg.values()) / 1e6 / steps, peak))
print('# %-78s %6s %10s %6s %9s %9s %8s %6s' % ('kernel', 'n/step', 'us/step', 'share', 'rd MB', 'wr MB', 'GB/s', 'of pk'))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / a[1] / 1e9 if a[1] else 0.0
    print('%-80s %6.1f %10.1f %5.1f%% %9.1f %9.1f %8.0f %5.0f%%' % (k[:80], a[0] / steps, a[1] * 1e6 / steps, 100 * a[1] / tt, a[2] / 1e6 / steps, a[3] / 1e6 / steps,
                                                               gbs, 100 * gbs / peak))
